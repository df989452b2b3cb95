// sparc_core.cuh -- per-env dynamics, rule validator and observation encoder of the batched
// SPaRC environment.  Everything here is written against explicit (thread id, thread count)
// arguments and a handful of wrapped intrinsics so that the *same source* compiles
//   * with nvcc for sm_100a (the product: kernels in sparc_b200.cu), and
//   * with g++ as plain C++ (tests/emu: a test-only harness that replays the kernels' logic on
//     the CPU so that logic bugs are found without a GPU).  The product never uses that build.
//
// Reference being restated: SPaRC_Gym/SPaRC_Gym.py ("SG.py"); SURVEY.md Appendix A is the spec.
#pragma once
#include <stdint.h>
#include "../../include/sparc_b200.h"

#if defined(SPARC_EMU_COUNTERS)
static long long g_fit_iters = 0, g_fit_calls = 0, g_fit_max = 0, g_fit_cur = 0;   // search effort of exact_fit (tests only)
static long long g_fit_hist[24];                                                   // searches by floor(log2(nodes)) (a call is booked when the next one begins)
static inline void sparc_fit_book() { if (g_fit_cur > 0) { int b = 0; for (long long v = g_fit_cur; v > 1; v >>= 1) b++; g_fit_hist[b < 23 ? b : 23]++; } }
#define SPARC_COUNT_FIT() (++g_fit_iters, ++g_fit_cur, g_fit_max = g_fit_cur > g_fit_max ? g_fit_cur : g_fit_max)
#define SPARC_COUNT_FIT_CALL() (sparc_fit_book(), ++g_fit_calls, g_fit_cur = 0)
#else
#define SPARC_COUNT_FIT() ((void)0)
#define SPARC_COUNT_FIT_CALL() ((void)0)
#endif

// -DVAL_PROFILE (tools/valprof.py): cycle counters of validate_env per queued terminal state -- where the ~30 us of
// k_val's slowest snapshot go.  g_valprof = sum total, sum prologue, sum flood fill, sum exact fit, snapshots,
// max total, max flood fill, max exact fit.  Not compiled otherwise.
#if defined(VAL_PROFILE) && defined(__CUDACC__)
__device__ unsigned long long g_valprof[8];
#endif
#if defined(VAL_PROFILE) && defined(__CUDA_ARCH__)
#define VP_BEGIN() long long vp_t0 = clock64(), vp_pro = 0, vp_flood = 0, vp_fit = 0, vp_tmp = 0
#define VP_PROLOGUE_DONE() (vp_pro = clock64() - vp_t0)
#define VP_TIC() (vp_tmp = clock64())
#define VP_TOC(acc) ((acc) += clock64() - vp_tmp)
#define VP_END() do { const unsigned long long vp_tot = (unsigned long long)(clock64() - vp_t0); \
    atomicAdd(&g_valprof[0], vp_tot); atomicAdd(&g_valprof[1], (unsigned long long)vp_pro); \
    atomicAdd(&g_valprof[2], (unsigned long long)vp_flood); atomicAdd(&g_valprof[3], (unsigned long long)vp_fit); \
    atomicAdd(&g_valprof[4], 1ull); atomicMax(&g_valprof[5], vp_tot); atomicMax(&g_valprof[6], (unsigned long long)vp_flood); \
    atomicMax(&g_valprof[7], (unsigned long long)vp_fit); } while (0)
#else
#define VP_BEGIN() ((void)0)
#define VP_PROLOGUE_DONE() ((void)0)
#define VP_TIC() ((void)0)
#define VP_TOC(acc) ((void)0)
#define VP_END() ((void)0)
#endif

#if defined(__CUDACC__)
#define SPARC_HD __host__ __device__ __forceinline__
#define SPARC_D __device__ __forceinline__
#else
#define SPARC_HD inline
#define SPARC_D inline
#endif

// ---------------------------------------------------------------------------------------------
// intrinsics (device versions map to single SASS instructions; host versions are for the emu)
// ---------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
SPARC_D uint32_t sp_brev(uint32_t v) { return __brev(v); }
SPARC_D int sp_popc(uint32_t v) { return __popc(v); }
SPARC_D int sp_popcll(uint64_t v) { return __popcll(v); }
SPARC_D int sp_ffs(uint32_t v) { return __ffs(v); }
SPARC_D int sp_ffsll(uint64_t v) { return __ffsll((long long)v); }
SPARC_D uint32_t sp_funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) { return __funnelshift_r(lo, hi, sh); }
template <typename T> SPARC_D T sp_ldg(const T* p) { return __ldg(p); }
struct sp_u4 { uint32_t x, y, z, w; };
SPARC_D sp_u4 sp_ldg4(const uint32_t* p) { uint4 v = __ldg(reinterpret_cast<const uint4*>(p)); return {v.x, v.y, v.z, v.w}; }
SPARC_D uint32_t sp_umulhi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
// streaming 128-bit store (st.global.cs): the observation is written once and not re-read by this kernel
SPARC_D void sp_st4(int32_t* p, int32_t a, int32_t b, int32_t c, int32_t d) {
#if defined(SPARC_STORE_PLAIN)
    *reinterpret_cast<int4*>(p) = make_int4(a, b, c, d);
#else
    __stcs(reinterpret_cast<int4*>(p), make_int4(a, b, c, d));
#endif
}
#else
SPARC_HD uint32_t sp_brev(uint32_t v) {
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
    v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
    return (v >> 16) | (v << 16);
}
SPARC_HD int sp_popc(uint32_t v) { return __builtin_popcount(v); }
SPARC_HD int sp_popcll(uint64_t v) { return __builtin_popcountll(v); }
SPARC_HD int sp_ffs(uint32_t v) { return __builtin_ffs((int)v); }
SPARC_HD int sp_ffsll(uint64_t v) { return __builtin_ffsll((long long)v); }
SPARC_HD uint32_t sp_funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {
    sh &= 31; return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
}
SPARC_HD uint32_t sp_umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
template <typename T> SPARC_HD T sp_ldg(const T* p) { return *p; }
struct sp_u4 { uint32_t x, y, z, w; };
SPARC_HD sp_u4 sp_ldg4(const uint32_t* p) { return {p[0], p[1], p[2], p[3]}; }
SPARC_HD void sp_st4(int32_t* p, int32_t a, int32_t b, int32_t c, int32_t d) { p[0] = a; p[1] = b; p[2] = c; p[3] = d; }
#endif

// ---------------------------------------------------------------------------------------------
// device-side views
// ---------------------------------------------------------------------------------------------
struct StoreDev {
    int n_puzzles, SX, SY, S;      // slab lattice, S = SX*SY nodes
    int P, L, W;                   // planes, L = P*S bits per env, W = ceil(L/32) words
    int nv;                        // words of one dense plane: ceil(S/32)
    int Wc;                        // words of the 4-bit colour string: ceil(4S/32)
    uint32_t inv_L, inv_S, inv_4S; // floor(2^32/d)+1: exact x/d by one mul-hi for the x used here
    uint32_t inv_k0b, inv_k0c, inv_k0z;   // same for the whole words of one env's strings: L/32, 4S/32, S/32 (0 when there is none)
    const SparcPuzzleHeader* hdr;
    const uint32_t* nb;            // [n][SPARC_N_NODE_BOARDS][8]
    const uint64_t* cb;            // [n][SPARC_N_CELL_BOARDS]
    const uint32_t* sbits;         // [n][W]
    const uint32_t* color4;        // [n][Wc]  colour ids, 4 bits per node, dense slab order
    const uint32_t* info_nz;       // [n][nv]  additional_info != 0
    const int32_t* info;           // [n][S]
    const SparcShape* shapes;
    const uint32_t* trie;          // [nodes][8]
    const uint32_t* vrec;          // [n][8][24] validator records of the warp-cooperative k_val (sparc_val_warp.cuh); device builds only
};

// per-env state, struct of arrays over envs (coalesced: consecutive lanes = consecutive envs)
struct StateDev {
    uint32_t* core;                // [N][4]: puzzle, step, trie node, x | y<<8 | on_depth<<16 | flags<<24
    uint32_t* vis;                 // [8][N]: visited plane, dense slab order (== set of path nodes)
    int32_t* ep_ret;               // [N]: episode return in cents
    uint32_t* moves;               // [12][N]: stack of 2-bit moves (traceback only), else nullptr
    unsigned long long* stats;     // [SPARC_N_STATS]
};

enum { FL_DONE = 1, FL_TERM = 2, FL_TRUNC = 4, FL_WON = 8 /* autoreset CONTINUE only: the last final step was a win */ };
#define SPARC_MOVE_WORDS 12

struct EnvRegs {                   // one env's state in registers
    uint32_t puzzle, step, tnode;
    int x, y, on_depth, flags;
    uint32_t v[8];
    int ep_ret;
};

struct StepCfg {
    long long n_envs;
    int traceback, max_steps, autoreset, stride, validate_mode;
};

struct StepResult {
    int reward_code, terminated, truncated, legal_mask, rule_flags;
};

SPARC_HD uint32_t board_word(const uint32_t (&b)[8], int k) {
    uint32_t r = b[0];
    r = (k == 1) ? b[1] : r; r = (k == 2) ? b[2] : r; r = (k == 3) ? b[3] : r; r = (k == 4) ? b[4] : r;
    r = (k == 5) ? b[5] : r; r = (k == 6) ? b[6] : r; r = (k == 7) ? b[7] : r;
    return r;
}
SPARC_HD int board_bit(const uint32_t (&b)[8], int n) { return (board_word(b, n >> 5) >> (n & 31)) & 1; }
SPARC_HD void board_flip(uint32_t (&b)[8], int n) {
    const uint32_t m = 1u << (n & 31);
    const int k = n >> 5;
#pragma unroll
    for (int i = 0; i < 8; i++) b[i] ^= (i == k) ? m : 0u;
}
SPARC_HD int board_popc(const uint32_t (&b)[8]) {
    int c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) c += sp_popc(b[i]);
    return c;
}
SPARC_HD void load_board(uint32_t (&b)[8], const uint32_t* p) {
    sp_u4 a = sp_ldg4(p), c = sp_ldg4(p + 4);
    b[0] = a.x; b[1] = a.y; b[2] = a.z; b[3] = a.w; b[4] = c.x; b[5] = c.y; b[6] = c.z; b[7] = c.w;
}

struct HdrRegs { int X, Y, sx, sy, tx, ty, n_shapes, flags; uint32_t shape_off, trie_root; };
SPARC_HD HdrRegs load_hdr(const StoreDev& st, uint32_t puzzle) {
    sp_u4 a = sp_ldg4(reinterpret_cast<const uint32_t*>(st.hdr + puzzle));
    HdrRegs h;
    h.X = a.x & 0xff; h.Y = (a.x >> 8) & 0xff; h.sx = (a.x >> 16) & 0xff; h.sy = a.x >> 24;
    h.tx = a.y & 0xff; h.ty = (a.y >> 8) & 0xff; h.n_shapes = (a.y >> 16) & 0xff; h.flags = a.y >> 24;
    h.shape_off = a.z; h.trie_root = a.w;
    return h;
}

// action id -> (dx, dy): right, up, left, down on [x, y] (SG.py:12-25, 212-217)
SPARC_HD int dir_dx(int a) { return (a == 0) - (a == 2); }
SPARC_HD int dir_dy(int a) { return (a == 3) - (a == 1); }

// ---------------------------------------------------------------------------------------------
// state load / store
// ---------------------------------------------------------------------------------------------
SPARC_HD void load_env(EnvRegs& e, const StateDev& sd, const StoreDev& st, long long n, long long i) {
    const uint32_t* c = sd.core + 4 * i;
#if defined(__CUDA_ARCH__)
    uint4 q = *reinterpret_cast<const uint4*>(c);
    e.puzzle = q.x; e.step = q.y; e.tnode = q.z; uint32_t w = q.w;
#else
    e.puzzle = c[0]; e.step = c[1]; e.tnode = c[2]; uint32_t w = c[3];
#endif
    e.x = w & 0xff; e.y = (w >> 8) & 0xff; e.on_depth = (w >> 16) & 0xff; e.flags = w >> 24;
#pragma unroll
    for (int k = 0; k < 8; k++) e.v[k] = (k < st.nv) ? sd.vis[(long long)k * n + i] : 0u;
    e.ep_ret = sd.ep_ret[i];
}
SPARC_HD void store_env(const EnvRegs& e, const StateDev& sd, const StoreDev& st, long long n, long long i) {
    uint32_t* c = sd.core + 4 * i;
    const uint32_t w = (uint32_t)e.x | ((uint32_t)e.y << 8) | ((uint32_t)e.on_depth << 16) | ((uint32_t)e.flags << 24);
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<uint4*>(c) = make_uint4(e.puzzle, e.step, e.tnode, w);
#else
    c[0] = e.puzzle; c[1] = e.step; c[2] = e.tnode; c[3] = w;
#endif
#pragma unroll
    for (int k = 0; k < 8; k++) if (k < st.nv) sd.vis[(long long)k * n + i] = e.v[k];
    sd.ep_ret[i] = e.ep_ret;
}

// clean reset (SG.py:1092-1095, 173-187; SURVEY Q1 neutralised)
SPARC_HD void env_reset(EnvRegs& e, const StoreDev& st, uint32_t puzzle) {
    const HdrRegs h = load_hdr(st, puzzle);
    e.puzzle = puzzle; e.step = 0; e.x = h.sx; e.y = h.sy; e.flags = 0; e.ep_ret = 0;
    e.tnode = h.trie_root; e.on_depth = h.trie_root ? 1 : 0;
#pragma unroll
    for (int k = 0; k < 8; k++) e.v[k] = 0;
    board_flip(e.v, h.sx * st.SY + h.sy);
}

// ---------------------------------------------------------------------------------------------
// dynamics (SG.py:1024-1051 legality, 1111-1223 step)
// ---------------------------------------------------------------------------------------------
// Legality (SG.py:1024-1051) on ONE board: avail = walk & ~visited (in this puzzle's lattice, not a gap, not yet on the
// path).  A move is legal when its target is available, or -- traceback -- when it is path[-2] (node index pn, or -1)
// and that node is walkable (prev_ok; it is, unless the start node itself sits on a gap).  One dynamic bit test per
// direction instead of two: the 8-way word selects of board_bit were half of the dynamics' instructions
// (profiles/r2_kernels_ncu_config2.txt, tools/envprof.py).
SPARC_HD void board_avail(uint32_t (&avail)[8], const uint32_t (&walk)[8], const uint32_t (&v)[8]) {
#pragma unroll
    for (int k = 0; k < 8; k++) avail[k] = walk[k] & ~v[k];
}
// bit a of the result: action a is legal from (x, y)
SPARC_HD int legal_actions(const StoreDev& st, const uint32_t (&avail)[8], int x, int y, int pn, bool prev_ok) {
    int mask = 0;
#pragma unroll
    for (int a = 0; a < 4; a++) {
        const int nx = x + dir_dx(a), ny = y + dir_dy(a);
        if ((unsigned)nx >= (unsigned)st.SX || (unsigned)ny >= (unsigned)st.SY) continue;
        const int n = nx * st.SY + ny;
        if (board_bit(avail, n) || (n == pn && prev_ok)) mask |= 1 << a;
    }
    return mask;
}
// node index of path[-2] from the position and the move that led there (d = top of the move stack)
SPARC_HD int prev_node(const StoreDev& st, int x, int y, int d) { return (x - dir_dx(d)) * st.SY + (y - dir_dy(d)); }

// top of the move stack = the move that led to path[-1]; path length = plen
SPARC_HD int stack_top(const StateDev& sd, long long n, long long i, int plen) {
    const int idx = plen - 2;
    return (sd.moves[(long long)(idx >> 4) * n + i] >> ((idx & 15) * 2)) & 3;
}

SPARC_HD void env_step(EnvRegs& e, const StoreDev& st, const StateDev& sd, const StepCfg& cfg, long long i,
                       long long action, StepResult& r) {
    const HdrRegs h = load_hdr(st, e.puzzle);
    uint32_t walk[8], avail[8];
    load_board(walk, st.nb + ((size_t)e.puzzle * SPARC_N_NODE_BOARDS + SPARC_NB_WALK) * 8);
    int plen = board_popc(e.v);                            // path nodes are distinct, visited == set(path)
    const int tb = cfg.traceback && sd.moves != nullptr;
    int pn = -1;                                           // path[-2] as a node index
    if (tb && plen >= 2) pn = prev_node(st, e.x, e.y, stack_top(sd, cfg.n_envs, i, plen));
    bool prev_ok = pn >= 0 && board_bit(walk, pn);
    board_avail(avail, walk, e.v);

    e.step += 1;                                           // SG.py:1132 (illegal actions count too)
    int truncated = (int)e.step >= cfg.max_steps;          // SG.py:1134
    int moved = 0;
    if (action >= 0 && action < 4) {
        const int a = (int)action;
        const int nx = e.x + dir_dx(a), ny = e.y + dir_dy(a);
        const int n = nx * st.SY + ny;
        const bool inside = (unsigned)nx < (unsigned)st.SX && (unsigned)ny < (unsigned)st.SY;
        const bool fresh = inside && board_bit(avail, n);
        if (fresh) {                                       // fresh node (SG.py:1167-1188)
            board_flip(e.v, n);
            if (e.on_depth == plen) {
                const uint32_t child = sp_ldg(st.trie + (size_t)e.tnode * SPARC_TRIE_NODE_WORDS + a);
                if (child) { e.tnode = child; e.on_depth = plen + 1; }
            }
            if (tb) {                                      // push the move
                const int idx = plen - 1;
                uint32_t* wp = sd.moves + (long long)(idx >> 4) * cfg.n_envs + i;
                const int sh = (idx & 15) * 2;
                *wp = (*wp & ~(3u << sh)) | ((uint32_t)a << sh);
                pn = e.x * st.SY + e.y; prev_ok = board_bit(walk, pn);
            }
            plen += 1;
            e.x = nx; e.y = ny;
            moved = 1;
        } else if (inside && n == pn && prev_ok) {         // traceback undo (SG.py:1141-1166)
            board_flip(e.v, e.x * st.SY + e.y);
            if (e.on_depth == plen) {                      // cursor leaves the popped node
                e.tnode = sp_ldg(st.trie + (size_t)e.tnode * SPARC_TRIE_NODE_WORDS + 4);
                e.on_depth = plen - 1;
            }
            plen -= 1;
            pn = -1; prev_ok = false;
            if (plen >= 2) { pn = prev_node(st, nx, ny, stack_top(sd, cfg.n_envs, i, plen)); prev_ok = board_bit(walk, pn); }
            e.x = nx; e.y = ny;
            moved = 1;
        }
        if (moved) board_avail(avail, walk, e.v);
    }
    const int terminated = (e.x == h.tx && e.y == h.ty);   // SG.py:1192
    const int legal_after = legal_actions(st, avail, e.x, e.y, pn, prev_ok);
    if (legal_after == 0) truncated = 1;                   // SG.py:1195-1196
    if (terminated) truncated = 0;                         // SG.py:1198-1199
    const int on_trie = (e.on_depth == plen) && e.tnode != 0;
    int code = SPARC_R_ZERO;
    const bool keep_going = cfg.autoreset == SPARC_AUTORESET_CONTINUE;
    if (terminated || truncated) {                         // SG.py:1204-1213
        const int is_end = on_trie ? (int)sp_ldg(st.trie + (size_t)e.tnode * SPARC_TRIE_NODE_WORDS + 5) : 0;
        code = is_end ? SPARC_R_WIN : SPARC_R_LOSE;
        // stepping past the end (the reference has no guard, Q4): its outcome_reward is only cleared by a
        // non-final step, so once it is 1 a later final step that is not on a listed path leaves it alone and
        // returns the normal reward of that step, 0 (SG.py:1205-1213: `if self.outcome_reward != 1`)
        if (keep_going) { if (is_end) e.flags |= FL_WON; else if (e.flags & FL_WON) code = SPARC_R_ZERO; }
    } else {
        if (keep_going) e.flags &= ~FL_WON;
        if (moved && (h.flags & 1)) code = on_trie ? SPARC_R_ON_PATH : SPARC_R_OFF_PATH;   // SG.py:1214-1223
    }
    e.ep_ret += (code == SPARC_R_ON_PATH) - (code == SPARC_R_OFF_PATH) + 100 * ((code == SPARC_R_WIN) - (code == SPARC_R_LOSE));
    r.reward_code = code; r.terminated = terminated; r.truncated = truncated; r.legal_mask = legal_after;
}

// legal mask of the current state without stepping (reset / frozen envs)
SPARC_HD int env_legal_now(const EnvRegs& e, const StoreDev& st, const StateDev& sd, const StepCfg& cfg, long long i) {
    uint32_t walk[8], avail[8];
    load_board(walk, st.nb + ((size_t)e.puzzle * SPARC_N_NODE_BOARDS + SPARC_NB_WALK) * 8);
    const int plen = board_popc(e.v);
    int pn = -1;
    if (cfg.traceback && sd.moves != nullptr && plen >= 2) pn = prev_node(st, e.x, e.y, stack_top(sd, cfg.n_envs, i, plen));
    const bool prev_ok = pn >= 0 && board_bit(walk, pn);
    board_avail(avail, walk, e.v);
    return legal_actions(st, avail, e.x, e.y, pn, prev_ok);
}

// ---------------------------------------------------------------------------------------------
// rule validator (SG.py:374-950) on 16-stride node boards and uint64 rule-cell boards
// ---------------------------------------------------------------------------------------------
// fill f (subset of p) along y inside every 16-bit row of the word; p has bits 15 and 31 clear
SPARC_HD uint32_t row_fill(uint32_t p, uint32_t f) {
    const uint32_t up = ((p ^ (p + f)) & p) | f;
    const uint32_t pr = sp_brev(p), fr = sp_brev(up);
    const uint32_t dn = ((pr ^ (pr + fr)) & pr) | fr;
    return sp_brev(dn);
}

struct Cnt3 { uint64_t c0, c1, c2; };                      // bit-sliced per-cell demand counters (0..7)
SPARC_HD void cnt_add(Cnt3& c, uint64_t m) { uint64_t t = c.c0 & m; c.c0 ^= m; m = t; t = c.c1 & m; c.c1 ^= m; c.c2 ^= t; }
SPARC_HD void cnt_sub(Cnt3& c, uint64_t m) { uint64_t t = ~c.c0 & m; c.c0 ^= m; m = t; t = ~c.c1 & m; c.c1 ^= m; c.c2 ^= t; }

// SparcShape words: mask = cells (bits 0..55) | anchor_dy << 56 | flags << 59; anchors = cells of THIS puzzle's
// board on which the shape's anchor may sit with every cell in bounds (bits 0..55, SG.py:880-883) | area << 56.
// One 16-byte load fetches everything a placement needs.
#define SPARC_CELLS56 0x00FFFFFFFFFFFFFFull
struct ShapeW { uint64_t mask, anchors; };
SPARC_HD ShapeW load_shape(const SparcShape* s) {
#if defined(__CUDA_ARCH__)
    const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(s));
    return {v.x, v.y};
#else
    return {s->mask, s->anchors};
#endif
}
SPARC_HD int shp_flags(uint64_t mask_word) { return (int)(mask_word >> 59) & 7; }
SPARC_HD int shp_area(uint64_t anchors_word) { return (int)(anchors_word >> 56); }
// cells covered when the shape's anchor sits on `cell` (0 = not placeable there)
SPARC_HD uint64_t shape_at(const ShapeW& s, int cell) {
    if (!((s.anchors >> cell) & 1ull)) return 0;
    return (s.mask & SPARC_CELLS56) << (cell - ((int)(s.mask >> 56) & 7));
}
SPARC_HD uint64_t shape_place(const SparcShape& s, int cx, int cy) { return shape_at(load_shape(&s), cx * 8 + cy); }

// 32 small counters (0..255) in four registers: instances per shape slot
struct Cnt8 { uint64_t w[4]; };
SPARC_HD uint64_t cnt8_word(const Cnt8& c, int k) { uint64_t r = c.w[0]; r = k == 1 ? c.w[1] : r; r = k == 2 ? c.w[2] : r; r = k == 3 ? c.w[3] : r; return r; }
SPARC_HD int cnt8_get(const Cnt8& c, int s) { return (int)(cnt8_word(c, s >> 3) >> ((s & 7) * 8)) & 255; }
SPARC_HD void cnt8_add(Cnt8& c, int s, int d) {              // d = +1 / -1 (or a first value into a zero slot)
    const uint64_t v = (uint64_t)(long long)d << ((s & 7) * 8);
    const int k = s >> 3;
#pragma unroll
    for (int i = 0; i < 4; i++) c.w[i] += (i == k) ? v : 0ull;
}

// exact-fit existence in the REFERENCE's search order (SG.py:738-895): all ylops first (any
// rule-cell anchor, may leave the region, may stack), then polys anchored at the first
// still-demanded cell.  Exponential in the number of ylops (49^k anchor combinations), so the
// kernels use exact_fit() below; this version is kept as an independent cross-check for the
// tests (tests/emu).  n_poly / n_ylop: instances per shape slot inside this region.
SPARC_HD bool exact_fit_reference_order(uint64_t region, int w, int h, const SparcShape* shapes, int ns,
                        const uint8_t* n_poly, const uint8_t* n_ylop) {
    SPARC_COUNT_FIT_CALL();
    uint8_t rem[SPARC_MAX_SHAPES];
    uint8_t ylist[SPARC_MAX_YLOPS];
    uint8_t fr_item[64], fr_pos[64];
    int ny = 0;
    for (int s = 0; s < ns; s++) {
        const int empty = (shp_flags(shapes[s].mask) & 2) != 0;       // zero-area shapes never touch the grid
        rem[s] = empty ? 0 : n_poly[s];
        for (int k = 0; k < (empty ? 0 : (int)n_ylop[s]); k++) if (ny < SPARC_MAX_YLOPS) ylist[ny++] = (uint8_t)s;
    }
    Cnt3 c = {region, 0, 0};
    int level = 0, it = 0;
    for (;;) {
        SPARC_COUNT_FIT();
        bool found = false;
        if (level < ny) {                                   // SG.py:794-816
            const SparcShape& sh = shapes[ylist[level]];
            for (int a = it; a < 64; a++) {
                const int cx = a >> 3, cy = a & 7;
                if (cx >= w || cy >= h) continue;
                const uint64_t m = shape_place(sh, cx, cy);
                if (!m) continue;
                cnt_add(c, m); fr_pos[level] = (uint8_t)a; level++; it = 0; found = true; break;
            }
        } else {                                            // SG.py:818-853
            const uint64_t nz = c.c0 | c.c1 | c.c2;
            if (nz == 0) return true;
            const int t = sp_ffsll(nz) - 1;
            for (int s = it; s < ns; s++) {
                if (!rem[s]) continue;
                const uint64_t m = shape_place(shapes[s], t >> 3, t & 7);
                if (!m || (m & ~nz)) continue;              // a covered cell would go positive (SG.py:822)
                cnt_sub(c, m); rem[s]--; fr_item[level] = (uint8_t)s; fr_pos[level] = (uint8_t)t;
                level++; it = 0; found = true; break;
            }
        }
        if (found) continue;
        if (level == 0) return false;
        level--;                                            // backtrack
        if (level < ny) {
            const int a = fr_pos[level];
            cnt_sub(c, shape_place(shapes[ylist[level]], a >> 3, a & 7));
            it = a + 1;
        } else {
            const int s = fr_item[level], t = fr_pos[level];
            cnt_add(c, shape_place(shapes[s], t >> 3, t & 7)); rem[s]++;
            it = s + 1;
        }
    }
}

// signed 4-bit bit-sliced per-cell counters (two's complement, -8..7): net demand of a cell =
// [cell in region] + ylops covering it - polys covering it
struct Cnt4 { uint64_t b0, b1, b2, b3; };
SPARC_HD void cnt4_add(Cnt4& c, uint64_t m) {
    uint64_t t = c.b0 & m; c.b0 ^= m; m = t; t = c.b1 & m; c.b1 ^= m; m = t; t = c.b2 & m; c.b2 ^= m; c.b3 ^= t;
}
SPARC_HD void cnt4_sub(Cnt4& c, uint64_t m) {
    uint64_t t = ~c.b0 & m; c.b0 ^= m; m = t; t = ~c.b1 & m; c.b1 ^= m; m = t; t = ~c.b2 & m; c.b2 ^= m; c.b3 ^= t;
}
SPARC_HD int cnt4_get(const Cnt4& c, int bit) {
    return (int)((c.b0 >> bit) & 1) + 2 * (int)((c.b1 >> bit) & 1) + 4 * (int)((c.b2 >> bit) & 1) - 8 * (int)((c.b3 >> bit) & 1);
}
// cells whose value equals the small constant k (-8..7)
SPARC_HD uint64_t cnt4_eq(const Cnt4& c, int k) {
    const uint64_t z = 0, o = ~z;
    return ~(c.b0 ^ ((k & 1) ? o : z)) & ~(c.b1 ^ ((k & 2) ? o : z)) & ~(c.b2 ^ ((k & 4) ? o : z)) & ~(c.b3 ^ ((k & 8) ? o : z));
}

// exact-fit existence (SG.py:738-895).  The reference accepts iff there is a placement of every
// ylop (anchor on any rule cell, all cells inside the grid, may leave the region, may stack) and
// of every non-empty poly (translation only) such that on every cell
//        #polys covering - #ylops covering == [cell in region].
// (Its own search places all ylops first -- 49^k anchor combinations -- and then anchors polys
// at the first still-demanded cell; SG.py:794-853.)  The same existence question is answered
// here by ONE sweep over the cells in (x, y) order: every shape is anchored at its
// lexicographically first cell (SG.py:864-865), so when the sweep reaches cell c everything
// that can still touch c must be anchored AT c; we choose the ylops and polys anchored at c
// (canonical non-decreasing slot order) so that c ends at exactly zero, then move on.  A ylop
// is therefore only ever tried where polys can absorb it, and a dead cell is detected once,
// not once per ylop combination.
// State lives in registers: the per-slot instance counters are packed (Cnt8), `have_*` has a bit per slot
// with instances left so that only those are visited, and a candidate placement costs ONE 16-byte load
// (beside the encoder every load waits behind its store stream: the search went at ~0.75 us per node when
// each candidate read five shape fields and two counter bytes from memory).
//
// The search is written as init + one-node step so that a warp can run 32 copies of it on disjoint parts of
// the tree (k_val: exact_fit_warp in sparc_val_warp.cuh): the first `nsplit` BRANCHING nodes on the way down (nodes
// with two or more valid children) hand their children out by ordinal parity -- bit b of `code` at the b-th
// branching node -- so the 2^nsplit codes partition the leaves and their OR is the answer.  nsplit = 0 is the
// plain depth-first search.
struct FitState {
    Cnt4 c;                                                   // net demand per cell
    Cnt8 rem_p, rem_y;                                        // instances left per slot
    uint64_t grid;
    uint32_t have_p, have_y, mono_slots;
    int tot_p, tot_y, level, cell, q;
    int b;                                                    // branching nodes above the current node (capped use: b < nsplit)
    int ord;                                                  // ordinal of the next valid child of the current node
    bool all_connected, resumed, split;
    uint8_t fr_cell[128], fr_q[128], fr_ord[128];             // DFS edges: q < 2ns placement, 0xFF = advance; ordinal | split << 7
};

// -1 = search on, 0 = no fit, 1 = fits
SPARC_HD int fit_init(FitState& f, uint64_t region, int w, int h, const SparcShape* shapes, int ns,
                      const Cnt8& n_poly, const Cnt8& n_ylop) {
    SPARC_COUNT_FIT_CALL();
    f.rem_p = n_poly; f.rem_y = n_ylop;
    f.tot_p = 0; f.tot_y = 0; f.have_p = 0; f.have_y = 0; f.mono_slots = 0; f.all_connected = true;
    // isolated-cell prune: once no ylop is left, a demanded cell with no demanded 4-neighbour can only be
    // covered by a one-cell poly (every other shape is 4-connected and would need a neighbour)
    for (int s = 0; s < ns; s++) {
        const ShapeW sh = load_shape(shapes + s);
        const int fl = shp_flags(sh.mask);
        int np = cnt8_get(f.rem_p, s), ny = cnt8_get(f.rem_y, s);
        if (fl & 2) { cnt8_add(f.rem_p, s, -np); cnt8_add(f.rem_y, s, -ny); np = ny = 0; }   // zero-area shapes never touch the grid
        if ((np | ny) && !(sh.anchors & SPARC_CELLS56)) return 0;          // fits nowhere on this board
        f.tot_p += np; f.tot_y += ny;
        if (np) { f.have_p |= 1u << s; if (!(fl & 4)) f.all_connected = false; }
        if (ny) f.have_y |= 1u << s;
        if ((sh.mask & SPARC_CELLS56) == 1ull) f.mono_slots |= 1u << s;
    }
    if (f.tot_y > SPARC_MAX_YLOPS) return 0;                  // (ingest guarantees this never happens)
    f.c.b0 = region; f.c.b1 = 0; f.c.b2 = 0; f.c.b3 = 0;
    f.grid = 0;
    for (int cx = 0; cx < w; cx++) f.grid |= ((1ull << h) - 1ull) << (cx * 8);
    f.level = 0; f.cell = 0; f.q = 0; f.b = 0; f.ord = 0; f.resumed = false; f.split = false;
    if (f.tot_y == 0) f.cell = sp_ffsll(region) - 1;          // no ylops: only demanded cells ever matter
    return -1;
}

// candidate children of the current node: a ylop / a poly of slot s anchored at f.cell
SPARC_HD bool fit_ylop_ok(const FitState& f, const SparcShape* shapes, int s, uint64_t floor_cells, uint64_t& m) {
    m = shape_at(load_shape(shapes + s), f.cell);
    return m && !(floor_cells & ~m);                          // one ylop fewer can lift later cells: those at the floor would be lost
}
SPARC_HD bool fit_poly_ok(const FitState& f, const SparcShape* shapes, int s, uint64_t floor_cells, uint64_t& m) {
    m = shape_at(load_shape(shapes + s), f.cell);
    if (!m || (floor_cells & m)) return false;                // would sink below what ylops can lift
    if (f.tot_y == 0 && f.all_connected) {
        Cnt4 c = f.c;
        cnt4_sub(c, m);
        const uint64_t D = (c.b0 | c.b1 | c.b2) & ~c.b3;      // cells still demanded
        const uint64_t iso = D & ~((D << 1) | (D >> 1) | (D << 8) | (D >> 8));
        if (iso) {
            int monos = -(int)((f.mono_slots >> s) & 1u);
            for (uint32_t ms = f.mono_slots; ms; ms &= ms - 1) monos += cnt8_get(f.rem_p, sp_ffs(ms) - 1);
            if (sp_popcll(iso) > monos) return false;
        }
    }
    return true;
}

SPARC_HD int fit_step(FitState& f, const SparcShape* shapes, int ns, int code, int nsplit) {
    SPARC_COUNT_FIT();
    const int n = cnt4_get(f.c, f.cell);
    const uint64_t later = f.grid & ~((2ull << f.cell) - 1ull);                 // cells after `cell`
    const uint64_t floor_cells = cnt4_eq(f.c, -f.tot_y) & later;                // later cells already at what ylops can lift
    const uint64_t open = (f.c.b0 | f.c.b1 | f.c.b2 | f.c.b3) & later;          // later cells with non-zero demand
    const uint32_t ycand = (f.q < ns && n + 1 <= f.tot_p) ? f.have_y & (0xFFFFFFFFu << f.q) : 0u;   // ylops here (slots >= q)
    const int q2 = f.q > ns ? f.q - ns : 0;
    const uint32_t pcand = (n > 0 && q2 < 32) ? f.have_p & (0xFFFFFFFFu << q2) : 0u;               // polys here (slots >= q - ns)
    // settled cell with work left: with ylops still to place every later cell is a candidate anchor; without, only
    // a demanded cell can take a poly, so the advance jumps straight to the next one
    const bool adv = n == 0 && (f.tot_p + f.tot_y) != 0 && later && (f.tot_y > 0 || open);
    uint64_t m;
    if (!f.resumed) {                                                           // first arrival at this node
        f.ord = 0; f.split = false;
        if (f.b < nsplit) {
            int kids = adv ? 1 : 0;
            for (uint32_t cand = ycand; cand; cand &= cand - 1) kids += fit_ylop_ok(f, shapes, sp_ffs(cand) - 1, floor_cells, m) ? 1 : 0;
            for (uint32_t cand = pcand; cand; cand &= cand - 1) kids += fit_poly_ok(f, shapes, sp_ffs(cand) - 1, floor_cells, m) ? 1 : 0;
            f.split = kids >= 2;
        }
    }
    f.resumed = false;
    const int skip = f.split ? ((code >> f.b) & 1) : 0;                         // this copy takes ordinals of the other parity
    const uint8_t tag = (uint8_t)(f.split ? 0x80 : 0);
    for (uint32_t cand = ycand; cand; cand &= cand - 1) {                       // anchor a ylop here
        const int s = sp_ffs(cand) - 1;
        if (!fit_ylop_ok(f, shapes, s, floor_cells, m)) continue;
        if (f.split && ((f.ord ^ skip) & 1)) { f.ord++; continue; }
        cnt4_add(f.c, m); cnt8_add(f.rem_y, s, -1); f.tot_y--;
        if (!cnt8_get(f.rem_y, s)) f.have_y &= ~(1u << s);
        f.fr_cell[f.level] = (uint8_t)f.cell; f.fr_q[f.level] = (uint8_t)s; f.fr_ord[f.level] = (uint8_t)(f.ord | tag); f.level++;
        f.q = s; f.b += f.split ? 1 : 0;
        return -1;
    }
    for (uint32_t cand = pcand; cand; cand &= cand - 1) {                       // anchor a poly here
        const int s = sp_ffs(cand) - 1;
        if (!fit_poly_ok(f, shapes, s, floor_cells, m)) continue;
        if (f.split && ((f.ord ^ skip) & 1)) { f.ord++; continue; }
        cnt4_sub(f.c, m); cnt8_add(f.rem_p, s, -1); f.tot_p--;
        if (!cnt8_get(f.rem_p, s)) f.have_p &= ~(1u << s);
        f.fr_cell[f.level] = (uint8_t)f.cell; f.fr_q[f.level] = (uint8_t)(ns + s); f.fr_ord[f.level] = (uint8_t)(f.ord | tag); f.level++;
        f.q = ns + s; f.b += f.split ? 1 : 0;
        return -1;
    }
    if (n == 0 && f.tot_p + f.tot_y == 0 && !open) return 1;                    // nothing left and nothing demanded
    if (adv && !(f.split && ((f.ord ^ skip) & 1))) {
        f.fr_cell[f.level] = (uint8_t)f.cell; f.fr_q[f.level] = 0xFF; f.fr_ord[f.level] = (uint8_t)(f.ord | tag); f.level++;
        f.cell = sp_ffsll(f.tot_y > 0 ? later : open) - 1; f.q = 0; f.b += f.split ? 1 : 0;
        return -1;
    }
    for (;;) {                                                                  // backtrack
        if (f.level == 0) return 0;
        f.level--;
        f.cell = f.fr_cell[f.level];
        const int qq = f.fr_q[f.level], o = f.fr_ord[f.level];
        f.b -= o >> 7;
        if (qq == 0xFF) continue;                                               // advance was the last child
        if (qq < ns) { cnt4_sub(f.c, shape_at(load_shape(shapes + qq), f.cell)); cnt8_add(f.rem_y, qq, 1); f.tot_y++; f.have_y |= 1u << qq; }
        else { cnt4_add(f.c, shape_at(load_shape(shapes + qq - ns), f.cell)); cnt8_add(f.rem_p, qq - ns, 1); f.tot_p++; f.have_p |= 1u << (qq - ns); }
        f.q = qq + 1; f.ord = (o & 0x7f) + 1; f.split = (o >> 7) != 0; f.resumed = true;
        return -1;
    }
}

// signed 6-bit bit-sliced per-cell counters (-32..31) for puzzles with more than six ylop instances
struct Cnt6 { uint64_t b[6]; };
struct NarrowCnt {
    typedef Cnt4 C;
    enum { max_ylops = SPARC_MAX_YLOPS };
    static SPARC_HD C make(uint64_t region) { C c = {region, 0, 0, 0}; return c; }
    static SPARC_HD void add(C& c, uint64_t m) { cnt4_add(c, m); }
    static SPARC_HD void sub(C& c, uint64_t m) { cnt4_sub(c, m); }
    static SPARC_HD int get(const C& c, int bit) { return cnt4_get(c, bit); }
    static SPARC_HD uint64_t eq(const C& c, int k) { return cnt4_eq(c, k); }
    static SPARC_HD uint64_t positive(const C& c) { return (c.b0 | c.b1 | c.b2) & ~c.b3; }
    static SPARC_HD uint64_t nonzero(const C& c) { return c.b0 | c.b1 | c.b2 | c.b3; }
};
struct WideCnt {
    typedef Cnt6 C;
    enum { max_ylops = SPARC_MAX_YLOPS_WIDE };
    static SPARC_HD C make(uint64_t region) { C c = {{region, 0, 0, 0, 0, 0}}; return c; }
    static SPARC_HD void add(C& c, uint64_t m) {
#pragma unroll
        for (int i = 0; i < 6; i++) { const uint64_t t = c.b[i] & m; c.b[i] ^= m; m = t; }
    }
    static SPARC_HD void sub(C& c, uint64_t m) {
#pragma unroll
        for (int i = 0; i < 6; i++) { const uint64_t t = ~c.b[i] & m; c.b[i] ^= m; m = t; }
    }
    static SPARC_HD int get(const C& c, int bit) {
        int v = 0;
#pragma unroll
        for (int i = 0; i < 5; i++) v += (int)((c.b[i] >> bit) & 1) << i;
        return v - 32 * (int)((c.b[5] >> bit) & 1);
    }
    static SPARC_HD uint64_t eq(const C& c, int k) {
        uint64_t r = ~0ull;
#pragma unroll
        for (int i = 0; i < 6; i++) r &= ~(c.b[i] ^ (((k >> i) & 1) ? ~0ull : 0ull));
        return r;
    }
    static SPARC_HD uint64_t positive(const C& c) { return (c.b[0] | c.b[1] | c.b[2] | c.b[3] | c.b[4]) & ~c.b[5]; }
    static SPARC_HD uint64_t nonzero(const C& c) { return c.b[0] | c.b[1] | c.b[2] | c.b[3] | c.b[4] | c.b[5]; }
};

// The plain depth-first search with its whole state in registers (no ordinal / split bookkeeping): what one thread runs
// for one region in the fused kernel, in k_val and in sparc_validate.  Same tree, same order as fit_init + fit_step with
// nsplit = 0 (tests/test_exact_fit.py counts the nodes of both).  1 = fits, 0 = does not, -1 = `budget` nodes spent
// (budget <= 0: unlimited).
// CW = the per-cell demand counters: NarrowCnt (4 bit planes: up to SPARC_MAX_YLOPS = 6 ylop instances, every pool
// the generator or the tests build) or WideCnt (6 bit planes: up to SPARC_MAX_YLOPS_WIDE = 30, puzzles whose header
// carries flag bit 2 -- the reference has no limit, but its own search enumerates 49^k ylop anchors).
template <class CW>
SPARC_HD int exact_fit_budget_t(uint64_t region, int w, int h, const SparcShape* shapes, int ns,
                                const Cnt8& n_poly, const Cnt8& n_ylop, int budget) {
    SPARC_COUNT_FIT_CALL();
    Cnt8 rem_p = n_poly, rem_y = n_ylop;
    uint8_t fr_cell[128], fr_q[128];                          // DFS edges: q < 2ns placement, 0xFF = advance
    int tot_p = 0, tot_y = 0;
    uint32_t have_p = 0, have_y = 0, mono_slots = 0;
    bool all_connected = true;
    // isolated-cell prune: once no ylop is left, a demanded cell with no demanded 4-neighbour can only be
    // covered by a one-cell poly (every other shape is 4-connected and would need a neighbour)
    for (int s = 0; s < ns; s++) {
        const ShapeW sh = load_shape(shapes + s);
        const int fl = shp_flags(sh.mask);
        int np = cnt8_get(rem_p, s), ny = cnt8_get(rem_y, s);
        if (fl & 2) { cnt8_add(rem_p, s, -np); cnt8_add(rem_y, s, -ny); np = ny = 0; }   // zero-area shapes never touch the grid
        if ((np | ny) && !(sh.anchors & SPARC_CELLS56)) return 0;       // fits nowhere on this board
        tot_p += np; tot_y += ny;
        if (np) { have_p |= 1u << s; if (!(fl & 4)) all_connected = false; }
        if (ny) have_y |= 1u << s;
        if ((sh.mask & SPARC_CELLS56) == 1ull) mono_slots |= 1u << s;
    }
    if (tot_y > CW::max_ylops) return 0;                  // (ingest guarantees this never happens)
    typename CW::C c = CW::make(region);
    uint64_t grid = 0;
    for (int cx = 0; cx < w; cx++) grid |= ((1ull << h) - 1ull) << (cx * 8);
    int level = 0, cell = 0, q = 0;
    if (tot_y == 0) cell = sp_ffsll(region) - 1;              // no ylops: only demanded cells ever matter
    for (int it = 0;; it++) {
        if (budget > 0 && it >= budget) return -1;
        SPARC_COUNT_FIT();
        const int n = CW::get(c, cell);
        const uint64_t later = grid & ~((2ull << cell) - 1ull);               // cells after `cell`
        const uint64_t floor_cells = CW::eq(c, -tot_y) & later;              // later cells already at what ylops can lift
        bool found = false;
        if (q < ns && n + 1 <= tot_p) {                                       // anchor a ylop here (slots >= q)
            for (uint32_t cand = have_y & (0xFFFFFFFFu << q); cand; cand &= cand - 1) {
                const int s = sp_ffs(cand) - 1;
                const uint64_t m = shape_at(load_shape(shapes + s), cell);
                if (!m) continue;
                if (floor_cells & ~m) continue;               // one ylop fewer can lift later cells: those at the floor would be lost
                CW::add(c, m); cnt8_add(rem_y, s, -1); tot_y--;
                if (!cnt8_get(rem_y, s)) have_y &= ~(1u << s);
                fr_cell[level] = (uint8_t)cell; fr_q[level] = (uint8_t)s; level++;
                q = s; found = true;
                break;
            }
        }
        if (!found && n > 0) {                                                // anchor a poly here (slots >= q - ns)
            const int q2 = q > ns ? q - ns : 0;
            for (uint32_t cand = q2 < 32 ? have_p & (0xFFFFFFFFu << q2) : 0u; cand; cand &= cand - 1) {
                const int s = sp_ffs(cand) - 1;
                const uint64_t m = shape_at(load_shape(shapes + s), cell);
                if (!m) continue;
                if (floor_cells & m) continue;                // would sink below what ylops can lift
                CW::sub(c, m); cnt8_add(rem_p, s, -1); tot_p--;
                if (tot_y == 0 && all_connected) {
                    const uint64_t D = CW::positive(c);                        // cells still demanded
                    const uint64_t iso = D & ~((D << 1) | (D >> 1) | (D << 8) | (D >> 8));
                    if (iso) {
                        int monos = 0;
                        for (uint32_t ms = mono_slots; ms; ms &= ms - 1) monos += cnt8_get(rem_p, sp_ffs(ms) - 1);
                        if (sp_popcll(iso) > monos) { CW::add(c, m); cnt8_add(rem_p, s, 1); tot_p++; continue; }
                    }
                }
                if (!cnt8_get(rem_p, s)) have_p &= ~(1u << s);
                fr_cell[level] = (uint8_t)cell; fr_q[level] = (uint8_t)(ns + s); level++;
                q = ns + s; found = true;
                break;
            }
        }
        if (found) continue;
        if (n == 0) {                                                         // settled: advance
            const uint64_t open = CW::nonzero(c) & later;                      // later cells with non-zero demand
            if (tot_p + tot_y == 0) { if (!open) return 1; }               // nothing left and nothing demanded
            else if (later && (tot_y > 0 || open)) {
                // with ylops still to place every later cell is a candidate anchor; without, only a
                // demanded cell can take a poly, so jump straight to the next one
                fr_cell[level] = (uint8_t)cell; fr_q[level] = 0xFF; level++;
                cell = sp_ffsll(tot_y > 0 ? later : open) - 1; q = 0;
                continue;
            }
        }
        for (;;) {                                                            // backtrack
            if (level == 0) return 0;
            level--;
            cell = fr_cell[level];
            const int qq = fr_q[level];
            if (qq == 0xFF) continue;                                         // advance was the last child
            if (qq < ns) { CW::sub(c, shape_at(load_shape(shapes + qq), cell)); cnt8_add(rem_y, qq, 1); tot_y++; have_y |= 1u << qq; }
            else { CW::add(c, shape_at(load_shape(shapes + qq - ns), cell)); cnt8_add(rem_p, qq - ns, 1); tot_p++; have_p |= 1u << (qq - ns); }
            q = qq + 1;
            break;
        }
    }
}

SPARC_HD int exact_fit_budget(uint64_t region, int w, int h, const SparcShape* shapes, int ns,
                              const Cnt8& n_poly, const Cnt8& n_ylop, int budget, bool wide = false) {
    if (wide) return exact_fit_budget_t<WideCnt>(region, w, h, shapes, ns, n_poly, n_ylop, 0);
    return exact_fit_budget_t<NarrowCnt>(region, w, h, shapes, ns, n_poly, n_ylop, budget);
}

// one copy of the search: `code` / `nsplit` select its share of the tree (0, 0 = all of it)
SPARC_HD bool exact_fit(uint64_t region, int w, int h, const SparcShape* shapes, int ns,
                        const Cnt8& n_poly, const Cnt8& n_ylop, int code = 0, int nsplit = 0) {
    if (nsplit == 0) return exact_fit_budget(region, w, h, shapes, ns, n_poly, n_ylop, 0) != 0;
    FitState f;
    int r = fit_init(f, region, w, h, shapes, ns, n_poly, n_ylop);
    while (r < 0) r = fit_step(f, shapes, ns, code, nsplit);
    return r != 0;
}

// vd: the env's visited plane in dense slab order, 9 readable words (word 8 may be anything
// when S <= 256).  region_map: optional int8[SX*SY] (row-major x, y), pre-filled with -1.
// region_fit: optional int8[SPARC_MAX_REGIONS], pre-filled with SPARC_FIT_NONE: the poly/ylop verdict of every
// region (SG.py:674-702); asking for it makes the poly rule look at every region instead of stopping at the
// first failure.
// fit_budget > 0: an exact-fit search that needs more nodes makes the call return -1 (k_val: the snapshot becomes a
// hard item, see k_val_hard).
SPARC_HD int validate_env(const StoreDev& st, uint32_t puzzle, const uint32_t* vd, int ax, int ay,
                          int8_t* region_map, int8_t* region_fit = nullptr, int fit_budget = 0) {
    VP_BEGIN();
    const HdrRegs h = load_hdr(st, puzzle);
    const uint32_t* nbp = st.nb + (size_t)puzzle * SPARC_N_NODE_BOARDS * 8;
    // visited: dense -> 16-stride
    uint32_t V[8];
#pragma unroll
    for (int k = 0; k < 8; k++) V[k] = 0;
    const uint32_t rowmask = (1u << st.SY) - 1u;
    for (int x = 0; x < st.SX; x++) {
        const int pos = x * st.SY;
        const uint32_t row = sp_funnel_r(vd[pos >> 5], vd[(pos >> 5) + 1], pos & 31) & rowmask;
        const uint32_t sh = row << ((x & 1) * 16);
#pragma unroll
        for (int k = 0; k < 8; k++) V[k] |= (k == (x >> 1)) ? sh : 0u;
    }
    uint32_t G[8], C[8], VAL[8], T[8];
    load_board(G, nbp + SPARC_NB_GAPS16 * 8);
    load_board(C, nbp + SPARC_NB_CELL16 * 8);
    load_board(VAL, nbp + SPARC_NB_VALID16 * 8);
    int flags = 0;
    if (ax == h.tx && ay == h.ty) flags |= SPARC_RULE_REACHED_TARGET;            // SG.py:488
    flags |= SPARC_RULE_PATH_NOT_CROSSING;      // path nodes are distinct by construction (SG.py:498)
    {                                                                           // SG.py:508, 521
        uint32_t gv = 0, dm = 0;
        load_board(T, nbp + SPARC_NB_DOT16 * 8);
#pragma unroll
        for (int k = 0; k < 8; k++) { gv |= G[k] & V[k]; dm |= T[k] & ~V[k]; }
        if (!gv) flags |= SPARC_RULE_NO_GAP_VIOLATIONS;
        if (!dm) flags |= SPARC_RULE_ALL_DOTS;
    }
    {   // triangles (SG.py:617-646): number of the 4 neighbour nodes on the path == count
        uint32_t bad = 0;
        uint32_t T1[8], T2[8], T3[8], T4[8];
        load_board(T1, nbp + SPARC_NB_TRI1 * 8); load_board(T2, nbp + SPARC_NB_TRI2 * 8);
        load_board(T3, nbp + SPARC_NB_TRI3 * 8); load_board(T4, nbp + SPARC_NB_TRI4 * 8);
        load_board(T, nbp + SPARC_NB_TRIBAD * 8);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t a = (V[k] << 16) | (k > 0 ? V[k - 1] >> 16 : 0u);   // node (x-1, y)
            const uint32_t b = (V[k] >> 16) | (k < 7 ? V[k + 1] << 16 : 0u);   // node (x+1, y)
            const uint32_t c = (V[k] << 1) & 0xFFFEFFFEu;                      // node (x, y-1)
            const uint32_t d = (V[k] >> 1) & 0x7FFF7FFFu;                      // node (x, y+1)
            const uint32_t s1 = a ^ b, c1 = a & b, s2 = c ^ d, c2 = c & d;
            const uint32_t b0 = s1 ^ s2, cr = s1 & s2, b1 = c1 ^ c2 ^ cr, b2 = c1 & c2;
            const uint32_t eq1 = b0 & ~b1 & ~b2, eq2 = ~b0 & b1 & ~b2, eq3 = b0 & b1, eq4 = b2;
            bad |= (T1[k] & ~eq1) | (T2[k] & ~eq2) | (T3[k] & ~eq3) | (T4[k] & ~eq4) | T[k];
        }
        if (!bad) flags |= SPARC_RULE_TRIANGLES;
    }
    // regions (SG.py:422-454): passable = rule cell, or neither gap nor path (Q6).  The three
    // region rules are vacuously true for a puzzle without stars, squares and shape instances,
    // so its region pass is skipped unless the caller wants the region map itself.
    if (!(h.flags & 2) && !region_map && !region_fit) {
        flags |= SPARC_RULE_SQUARES | SPARC_RULE_STARS | SPARC_RULE_POLY_YLOP;
        if ((flags & 0xFF) == 0xFF) flags |= SPARC_RULE_ALL;
        VP_PROLOGUE_DONE();
        VP_END();
        return flags;
    }
    uint32_t Pm[8], unl[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { Pm[k] = VAL[k] & (C[k] | ~(G[k] | V[k])); unl[k] = C[k]; }
    const uint64_t* cbp = st.cb + (size_t)puzzle * SPARC_N_CELL_BOARDS;
    const uint64_t STAR = sp_ldg(cbp + SPARC_CB_STAR), SQ = sp_ldg(cbp + SPARC_CB_SQUARE);
    const uint64_t SB0 = sp_ldg(cbp + SPARC_CB_SYM1), SB1 = sp_ldg(cbp + SPARC_CB_SYM2);
    uint64_t COL[8], anycol = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) { COL[c] = sp_ldg(cbp + SPARC_CB_COLOR1 + c); anycol |= COL[c]; }
    const SparcShape* shapes = st.shapes + h.shape_off;
    const int w = (h.X - 1) >> 1, hh = (h.Y - 1) >> 1;
    bool sq_ok = true, star_ok = true, poly_ok = true;
    int rid = 0;
    VP_PROLOGUE_DONE();
    for (;;) {
        int k0 = -1;
#pragma unroll
        for (int k = 7; k >= 0; k--) k0 = unl[k] ? k : k0;
        if (k0 < 0) break;
        uint32_t f[8];
        const uint32_t seed = board_word(unl, k0) & (0u - board_word(unl, k0));    // lowest set bit: raster order
#pragma unroll
        for (int k = 0; k < 8; k++) f[k] = (k == k0) ? seed : 0u;
        VP_TIC();
        for (;;) {                                                              // flood to fix-point
            uint32_t g[8], ch = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) g[k] = row_fill(Pm[k], f[k]);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint32_t up = (g[k] << 16) | (k > 0 ? g[k - 1] >> 16 : 0u);
                const uint32_t dn = (g[k] >> 16) | (k < 7 ? g[k + 1] << 16 : 0u);
                const uint32_t nw = (g[k] | up | dn) & Pm[k];
                ch |= nw ^ f[k];
                f[k] = nw;
            }
            if (!ch) break;
        }
        VP_TOC(vp_flood);
        uint64_t R = 0;                                     // component -> rule-cell board
#pragma unroll
        for (int cx = 0; cx < 7; cx++) {
            const uint32_t cells = f[cx] & C[cx];
            unl[cx] &= ~cells;
            uint32_t t = (cells >> 17) & 0x5555u;           // x = 2cx+1 is the high half; y odd -> even slots
            t = (t | (t >> 1)) & 0x3333u; t = (t | (t >> 2)) & 0x0F0Fu; t = (t | (t >> 4)) & 0x00FFu;
            R |= (uint64_t)t << (cx * 8);
        }
        unl[7] &= ~(f[7] & C[7]);
        if (region_map) {
            uint64_t m = R;
            while (m) { const int b = sp_ffsll(m) - 1; m &= m - 1; region_map[(2 * (b >> 3) + 1) * st.SY + 2 * (b & 7) + 1] = (int8_t)rid; }
        }
        // squares (SG.py:534-552): at most one distinct non-zero colour per region
        const uint64_t sq = SQ & R;
        if (sq) {
            int nc = 0;
#pragma unroll
            for (int c = 0; c < 8; c++) nc += (sq & COL[c]) != 0;
            if (nc > 1) sq_ok = false;
        }
        // stars (SG.py:554-614): every star colour must total exactly 2 symbols in the region
        const uint64_t sr = STAR & R;
        if (sr) {
            if (sr & ~anycol) star_ok = false;              // colourless star (SG.py:580-582)
#pragma unroll
            for (int c = 0; c < 8; c++) {
                if (sr & COL[c]) {
                    const uint64_t rc = R & COL[c];
                    if (sp_popcll(SB0 & rc) + 2 * sp_popcll(SB1 & rc) != 2) star_ok = false;
                }
            }
        }
        // polys / ylops (SG.py:648-710)
        if (h.n_shapes && (poly_ok || region_fit)) {
            Cnt8 np = {{0, 0, 0, 0}}, nyl = {{0, 0, 0, 0}};
            int net = 0, any = 0;
            for (int s = 0; s < h.n_shapes; s++) {
                const int a = sp_popcll(sp_ldg(&shapes[s].poly_cells) & R), b = sp_popcll(sp_ldg(&shapes[s].ylop_cells) & R);
                if (a | b) {
                    cnt8_add(np, s, a); cnt8_add(nyl, s, b); any = 1;
                    net += (a - b) * shp_area(sp_ldg(&shapes[s].anchors));
                }
            }
            if (any) {
                int verdict = SPARC_FIT_OK;
                VP_TIC();
                if (net != sp_popcll(R)) verdict = SPARC_FIT_AREA_MISMATCH;   // SG.py:679-682
                else {
                    const int fit = exact_fit_budget(R, w, hh, shapes, h.n_shapes, np, nyl, fit_budget, (h.flags & 4) != 0);
                    if (fit < 0) return -1;
                    if (!fit) verdict = SPARC_FIT_NO_TILING;
                }
                VP_TOC(vp_fit);
                if (verdict != SPARC_FIT_OK) poly_ok = false;
                if (region_fit && rid < SPARC_MAX_REGIONS) region_fit[rid] = (int8_t)verdict;
            }
        }
        rid++;
    }
    if (sq_ok) flags |= SPARC_RULE_SQUARES;
    if (star_ok) flags |= SPARC_RULE_STARS;
    if (poly_ok) flags |= SPARC_RULE_POLY_YLOP;
    if ((flags & 0xFF) == 0xFF) flags |= SPARC_RULE_ALL;                        // SG.py:932-935
    VP_END();
    return flags;
}

// ---------------------------------------------------------------------------------------------
// observation encoder (SG.py:956-979 with a full rewrite in the reference dtypes)
//
//   tile : nE consecutive envs starting at env0 (env0 % 4 == 0 keeps every stream 16-B aligned).
//   The tile's output is three contiguous streams (base int32[nE][P][SX][SY], colour
//   int32[nE][SX][SY], additional_info int64[nE][SX][SY]); each is written in 16-byte chunks,
//   chunk c by thread c mod nthr, so a warp store covers 512 contiguous bytes.
//
//   Shared-memory staging ("rows"): for each stream, env j's bit string (L bits: 1 bit per base
//   element, 4 per colour element, 1 "non-zero" bit per info element) is stored in its own row
//   of W+2 words.  Bit g of the tile-wide stream (g = j*L + r) sits at row-bit r of row j: two
//   LDS + one funnel shift fetch the 32 bits from there on.  The few bits of env j+1 that share
//   env j's last chunk are appended to row j ("tail").
//   dyn : SPARC_DYN_STRIDE words per env: visited plane (dense, 8 words) + bit index of the agent
// ---------------------------------------------------------------------------------------------
#define SPARC_DYN_STRIDE 9
#define SPARC_VQ_WORDS 12     // a state to validate: env, puzzle, x | y << 8, counts-for-RULES_PASS, visited[8]

SPARC_HD void write_dyn(const StoreDev& st, uint32_t* dyn, int j, const EnvRegs& e) {
    uint32_t* row = dyn + j * SPARC_DYN_STRIDE;
#pragma unroll
    for (int k = 0; k < 8; k++) row[k] = e.v[k];
    row[8] = (uint32_t)(2 * st.S + e.x * st.SY + e.y);       // agent_location is plane 2
}

// pull one puzzle's encoder inputs (plane bits, colour nibbles, info mask) towards this SM's L1
SPARC_HD void prefetch_static(const StoreDev& st, uint32_t puzzle) {
#if defined(__CUDA_ARCH__)
    const char* a = reinterpret_cast<const char*>(st.sbits + (size_t)puzzle * st.W);
    for (int off = 0; off < st.W * 4; off += 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(a + off));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(a + st.W * 4 - 4));
    const char* c = reinterpret_cast<const char*>(st.color4 + (size_t)puzzle * st.Wc);
    asm volatile("prefetch.global.L1 [%0];" ::"l"(c));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(c + st.Wc * 4 - 4));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(st.info_nz + (size_t)puzzle * st.nv));
#else
    (void)st; (void)puzzle;
#endif
}

// bit-string sources: word k of env j's string (0 outside [0, W))
struct PlaneSrc {                       // the P one-hot planes: static planes of the puzzle | visited (plane 0) | agent bit
    const StoreDev& st; const uint32_t* dyn; const uint32_t* tile_puzzle;
    SPARC_HD int L() const { return st.L; }
    SPARC_HD int W() const { return st.W; }
    // (32-bit word offsets into the store arrays: n_puzzles * W < 2^32 is checked by sparc_store_create)
    SPARC_HD uint32_t raw(int j, int k) const {                 // the global-memory part (issued early, in batches)
        return (unsigned)k < (unsigned)st.W ? sp_ldg(st.sbits + (tile_puzzle[j] * (uint32_t)st.W + (uint32_t)k)) : 0u;
    }
    SPARC_HD uint32_t fix(int j, int k, uint32_t w) const {     // the shared-memory part (k >= 0; words >= W stay 0)
        const uint32_t* row = dyn + j * SPARC_DYN_STRIDE;
        const uint32_t v = (unsigned)k < (unsigned)st.nv ? row[k] : 0u;       // visited = plane 0
        const uint32_t ab = row[8];                                           // agent bit (inside plane 2)
        return w | v | ((int)(ab >> 5) == k ? 1u << (ab & 31) : 0u);
    }
    SPARC_HD uint32_t word(int j, int k) const { return fix(j, k, raw(j, k)); }
};
struct ColorSrc {                       // colour plane, 4 bits per node
    const StoreDev& st; const uint32_t* tile_puzzle;
    SPARC_HD int L() const { return 4 * st.S; }
    SPARC_HD int W() const { return st.Wc; }
    SPARC_HD uint32_t raw(int j, int k) const {
        return (unsigned)k < (unsigned)st.Wc ? sp_ldg(st.color4 + (tile_puzzle[j] * (uint32_t)st.Wc + (uint32_t)k)) : 0u;
    }
    SPARC_HD uint32_t fix(int, int, uint32_t w) const { return w; }
    SPARC_HD uint32_t word(int j, int k) const { return raw(j, k); }
};
struct InfoNzSrc {                      // additional_info != 0, 1 bit per node
    const StoreDev& st; const uint32_t* tile_puzzle;
    SPARC_HD int L() const { return st.S; }
    SPARC_HD int W() const { return st.nv; }
    SPARC_HD uint32_t raw(int j, int k) const {
        return (unsigned)k < (unsigned)st.nv ? sp_ldg(st.info_nz + (tile_puzzle[j] * (uint32_t)st.nv + (uint32_t)k)) : 0u;
    }
    SPARC_HD uint32_t fix(int, int, uint32_t w) const { return w; }
    SPARC_HD uint32_t word(int j, int k) const { return raw(j, k); }
};

// stage 1 (cooperative): rows[j][k] = word k of env j's string, k in [0, W+1] (zeros behind the string), and -- in the
// words from k0 = L/32 on -- the first bits of env j+1 that belong to env j's last chunk (the "tail"; A = stream bits
// per 16-byte chunk, a power of two <= 32).  Two parts, no barrier between them: the whole words [0, k0) of every row,
// four per thread per pass with all global loads issued before any is used; then the two or three last words of row j
// by one thread (counted from the end of the CTA: the first threads ran phase A in the fused kernel), which also
// fetches word 0 of env j+1.  (Rows used to be stored shifted by (j*L) mod 32 so that a chunk was one LDS + one shift;
// every row word then cost two loads, two fix-ups and a funnel shift -- ~80 instructions, a quarter of
// k_encode<reference> and half of k_encode<compact>.  Plain rows cost ~20 per word and one more LDS + a funnel shift
// per fetch.  The tails used to be a second stage behind a barrier: ~450 instructions in ONE warp, 0.7 us of the
// small-batch kernel's critical path -- tools/envprof.py; folded into every word of the first part they cost the
// issue-bound compact encoder 7 us of 60.)
template <class Src>
SPARC_HD void build_rows(const Src& src, uint32_t* rows, int nE, uint32_t inv_k0, int A, int tid, int nthr) {
    const int L = src.L(), Ws = src.W() + 2;
    const int k0 = L >> 5, sh = L & 31;                         // row-bit L: right behind env j's own bits
    const int total = nE * k0;
    for (int base = tid; base < total; base += 4 * nthr) {
        int j[4], k[4];
        uint32_t w[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int idx = base + u * nthr;
            j[u] = k0 == 1 ? idx : (int)sp_umulhi((uint32_t)idx, inv_k0); k[u] = idx - j[u] * k0;   // (2^32 / 1 + 1 does not fit)
            w[u] = idx < total ? src.raw(j[u], k[u]) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (base + u * nthr < total) rows[j[u] * Ws + k[u]] = src.fix(j[u], k[u], w[u]);
    }
    for (int j = nthr - 1 - tid; j < nE; j += nthr) {
        const bool more = j + 1 < nE;
        const uint32_t w0 = src.raw(j, k0), w1 = src.raw(j, k0 + 1), w2 = src.raw(j, k0 + 2), t = more ? src.raw(j + 1, 0) : 0u;
        const int tb = (A - (((j + 1) * L) & (A - 1))) & (A - 1);             // bits of env j+1 in env j's last chunk
        const uint32_t bits = more ? src.fix(j + 1, 0, t) & ((1u << tb) - 1u) : 0u;
        uint32_t* r = rows + j * Ws + k0;
        r[0] = src.fix(j, k0, w0) | (bits << sh);
        r[1] = src.fix(j, k0 + 1, w1) | (sh ? bits >> (32 - sh) : 0u);
        if (k0 + 2 < Ws) r[2] = src.fix(j, k0 + 2, w2);
    }
}
// stage 2 (cooperative): expand.  g = first stream bit of the chunk -> (env j, row bit); returns the 32 stream
// bits from g on (bits behind the row's tail are zero / unspecified)
SPARC_HD uint32_t row_fetch(const uint32_t* rows, uint32_t g, int L, int Ws, uint32_t inv_L) {
    const int j = (int)sp_umulhi(g, inv_L);
    const uint32_t pos = g - (uint32_t)(j * L);
    const uint32_t* r = rows + j * Ws + (pos >> 5);
    return sp_funnel_r(r[0], r[1], pos & 31u);
}
SPARC_HD void expand_base(const StoreDev& st, const uint32_t* rows, int32_t* out_tile, int nE, int tid, int nthr) {
    const int L = st.L, Ws = st.W + 2, total = nE * L, nchunks = total >> 2;
    int c = tid;
    for (; c + 3 * nthr < nchunks; c += 4 * nthr) {             // 4 independent stores in flight per thread
        uint32_t n[4];
#pragma unroll
        for (int u = 0; u < 4; u++) n[u] = row_fetch(rows, 4u * (uint32_t)(c + u * nthr), L, Ws, st.inv_L);
#pragma unroll
        for (int u = 0; u < 4; u++)
            sp_st4(out_tile + 4 * (size_t)(c + u * nthr), n[u] & 1, (n[u] >> 1) & 1, (n[u] >> 2) & 1, (n[u] >> 3) & 1);
    }
    for (; c < nchunks; c += nthr) {
        const uint32_t n = row_fetch(rows, 4u * (uint32_t)c, L, Ws, st.inv_L);
        sp_st4(out_tile + 4 * (size_t)c, n & 1, (n >> 1) & 1, (n >> 2) & 1, (n >> 3) & 1);
    }
    for (int e = (nchunks << 2) + tid; e < total; e += nthr)    // tail (< 4 elements)
        out_tile[e] = row_fetch(rows, (uint32_t)e, L, Ws, st.inv_L) & 1u;
}
// colour plane: 4-bit colour ids -> int32 output (a chunk = 16 stream bits)
SPARC_HD void expand_color(const StoreDev& st, const uint32_t* rows, int32_t* out_tile, int nE, int tid, int nthr) {
    const int L = 4 * st.S, Ws = st.Wc + 2, total = nE * st.S, nchunks = total >> 2;
    for (int c = tid; c < nchunks; c += nthr) {
        const uint32_t n = row_fetch(rows, 16u * (uint32_t)c, L, Ws, st.inv_4S);
        sp_st4(out_tile + 4 * (size_t)c, n & 15, (n >> 4) & 15, (n >> 8) & 15, (n >> 12) & 15);
    }
    for (int e = (nchunks << 2) + tid; e < total; e += nthr)
        out_tile[e] = row_fetch(rows, 4u * (uint32_t)e, L, Ws, st.inv_4S) & 15u;
}
// additional_info plane -> int64 output (two elements per 16-byte store).  The plane is almost
// all zeros: the non-zero mask decides, and only marked nodes read their int32 value.
SPARC_HD void expand_info(const StoreDev& st, const uint32_t* rows, const uint32_t* tile_puzzle, int64_t* out_tile,
                          int nE, int tid, int nthr) {
    const int S = st.S, Ws = st.nv + 2, total = nE * S, nchunks = total >> 1;
    for (int c = tid; c < nchunks; c += nthr) {
        const uint32_t e0 = 2u * (uint32_t)c;
        const uint32_t m = row_fetch(rows, e0, S, Ws, st.inv_S) & 3u;
        int32_t a = 0, b = 0;
        if (m) {
            const int j = (int)sp_umulhi(e0, st.inv_S), r = (int)e0 - j * S;
            if (m & 1) a = sp_ldg(st.info + (size_t)tile_puzzle[j] * S + r);
            if (m & 2) {
                const int j2 = (r + 1 >= S) ? j + 1 : j, r2 = (r + 1 >= S) ? 0 : r + 1;
                b = sp_ldg(st.info + (size_t)tile_puzzle[j2] * S + r2);
            }
        }
        sp_st4(reinterpret_cast<int32_t*>(out_tile + 2 * (size_t)c), a, a >> 31, b, b >> 31);
    }
    if ((total & 1) && tid == 0) {
        const int e = total - 1, jj = e / S, rr = e - jj * S;
        out_tile[e] = sp_ldg(st.info + (size_t)tile_puzzle[jj] * S + rr);
    }
}

// ---- compact format: uint8 planes / uint8 colour / int32 additional_info (same values, narrowest types) ----
// 4 bits -> 4 bytes of 0/1 (bit i lands in byte i)
SPARC_HD uint32_t spread4(uint32_t nib) { return ((nib & 15u) * 0x00204081u) & 0x01010101u; }
// 4 nibbles (16 bits) -> 4 bytes
SPARC_HD uint32_t nib4(uint32_t v) { return (v & 0xFu) | ((v & 0xF0u) << 4) | ((v & 0xF00u) << 8) | ((v & 0xF000u) << 12); }
// planes: a 16-byte chunk = 16 elements = 16 stream bits
// (A 256 x 8-byte shared-memory table -- stream byte -> 8 output bytes, two LDS.64 per chunk instead of the four
// shift/mask/multiply/mask groups -- was measured slower: k_encode<compact> 59.7 -> 65.5 us; the loop is issue-bound
// but the table's bank conflicts cost more than the ALU work they replace.)
SPARC_HD void expand_base_u8(const StoreDev& st, const uint32_t* rows, uint8_t* out_tile, int nE, int tid, int nthr) {
    const int L = st.L, Ws = st.W + 2, total = nE * L, nchunks = total >> 4;
    for (int c = tid; c < nchunks; c += nthr) {
        const uint32_t n = row_fetch(rows, 16u * (uint32_t)c, L, Ws, st.inv_L);
        sp_st4(reinterpret_cast<int32_t*>(out_tile + 16 * (size_t)c), (int32_t)spread4(n), (int32_t)spread4(n >> 4),
               (int32_t)spread4(n >> 8), (int32_t)spread4(n >> 12));
    }
    for (int e = (nchunks << 4) + tid; e < total; e += nthr)
        out_tile[e] = (uint8_t)(row_fetch(rows, (uint32_t)e, L, Ws, st.inv_L) & 1u);
}
// colour: a chunk = 16 elements = 64 stream bits = two word-aligned 32-bit fetches (each may sit in another env)
SPARC_HD void expand_color_u8(const StoreDev& st, const uint32_t* rows, uint8_t* out_tile, int nE, int tid, int nthr) {
    const int L = 4 * st.S, Ws = st.Wc + 2, total = nE * st.S, nchunks = total >> 4;
    for (int c = tid; c < nchunks; c += nthr) {
        const uint32_t lo = row_fetch(rows, 64u * (uint32_t)c, L, Ws, st.inv_4S);
        const uint32_t hi = row_fetch(rows, 64u * (uint32_t)c + 32u, L, Ws, st.inv_4S);
        sp_st4(reinterpret_cast<int32_t*>(out_tile + 16 * (size_t)c), (int32_t)nib4(lo), (int32_t)nib4(lo >> 16),
               (int32_t)nib4(hi), (int32_t)nib4(hi >> 16));
    }
    for (int e = (nchunks << 4) + tid; e < total; e += nthr)
        out_tile[e] = (uint8_t)(row_fetch(rows, 4u * (uint32_t)e, L, Ws, st.inv_4S) & 15u);
}
// additional_info: a chunk = 4 int32 elements; only marked nodes read their value
SPARC_HD void expand_info_i32(const StoreDev& st, const uint32_t* rows, const uint32_t* tile_puzzle, int32_t* out_tile,
                              int nE, int tid, int nthr) {
    const int S = st.S, Ws = st.nv + 2, total = nE * S, nchunks = total >> 2;
    for (int c = tid; c < nchunks; c += nthr) {
        const uint32_t e0 = 4u * (uint32_t)c;
        const uint32_t m = row_fetch(rows, e0, S, Ws, st.inv_S) & 15u;
        int32_t v[4] = {0, 0, 0, 0};
        if (m) {
            int j = (int)sp_umulhi(e0, st.inv_S), r = (int)e0 - j * S;
            if (r + 3 < S) {                                     // the four elements belong to one env: one base address
                const int32_t* ip = st.info + ((size_t)tile_puzzle[j] * S + r);
#pragma unroll
                for (int u = 0; u < 4; u++) if ((m >> u) & 1u) v[u] = sp_ldg(ip + u);
            } else {
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if ((m >> u) & 1u) v[u] = sp_ldg(st.info + (size_t)tile_puzzle[j] * S + r);
                    if (++r >= S) { r = 0; j++; }
                }
            }
        }
        sp_st4(out_tile + 4 * (size_t)c, v[0], v[1], v[2], v[3]);
    }
    for (int e = (nchunks << 2) + tid; e < total; e += nthr) {
        const int jj = e / S, rr = e - jj * S;
        out_tile[e] = sp_ldg(st.info + (size_t)tile_puzzle[jj] * S + rr);
    }
}

template <int FMT> struct ObsTypes { typedef int32_t base_t; typedef int64_t info_t; };          // SPARC_OBS_REFERENCE
template <> struct ObsTypes<SPARC_OBS_COMPACT> { typedef uint8_t base_t; typedef int32_t info_t; };

// shared-memory words: per-tile header (dyn rows + puzzle ids) and the three row sets
SPARC_HD int tile_head_words(int tile) { return tile * (SPARC_DYN_STRIDE + 1); }
// the fused k_env also keeps, behind the rows: per-env flags of phase A (PA_*) and the snapshots its validator warp works on
SPARC_HD int tile_snap_words(int tile) { return tile * (1 + SPARC_VQ_WORDS); }
SPARC_HD int tile_flat_words(const StoreDev& st, int tile) { return tile * ((st.W + 2) + (st.Wc + 2) + (st.nv + 2)); }
// encoder for one tile: stage 1, barrier, stage 2 (barrier among the nthr encoding threads).
// head = dyn[tile][9] | tile_puzzle[tile]; flats = base rows | colour rows | info rows
// FMT = SPARC_OBS_* is a compile-time constant in the kernels: the reference-format encoder carries no trace of the
// compact one (a run-time branch here cost 4 % of k_encode: 0.170 -> 0.176 ms, measured)
template <int FMT = SPARC_OBS_REFERENCE>
SPARC_HD void encode_tile_stage1(const StoreDev& st, const uint32_t* head, uint32_t* flats, int tile, int nE,
                                 const SparcObsOut& obs, int tid, int nthr) {
    const uint32_t* dyn = head;
    const uint32_t* tile_puzzle = head + tile * SPARC_DYN_STRIDE;
    uint32_t* rb = flats;
    uint32_t* rc = rb + tile * (st.W + 2);
    uint32_t* rz = rc + tile * (st.Wc + 2);
    // stream bits per fetch of the expand loops: 16-byte chunks of int32 / int32 / int64, or (compact) uint8 /
    // uint8 fetched as two 32-bit halves / int32
    constexpr bool cp = FMT == SPARC_OBS_COMPACT;
    if (obs.base_dev) build_rows(PlaneSrc{st, dyn, tile_puzzle}, rb, nE, st.inv_k0b, cp ? 16 : 4, tid, nthr);
    if (obs.color_dev) build_rows(ColorSrc{st, tile_puzzle}, rc, nE, st.inv_k0c, cp ? 32 : 16, tid, nthr);
    if (obs.info_dev) build_rows(InfoNzSrc{st, tile_puzzle}, rz, nE, st.inv_k0z, cp ? 4 : 2, tid, nthr);
}
template <int FMT = SPARC_OBS_REFERENCE>
SPARC_HD void encode_tile_stage2(const StoreDev& st, const uint32_t* head, const uint32_t* flats, int tile, int nE,
                                 long long env0, const SparcObsOut& obs, int tid, int nthr) {
    const uint32_t* tile_puzzle = head + tile * SPARC_DYN_STRIDE;
    const uint32_t* rb = flats;
    const uint32_t* rc = rb + tile * (st.W + 2);
    const uint32_t* rz = rc + tile * (st.Wc + 2);
    if (FMT == SPARC_OBS_COMPACT) {              // (env0 is a multiple of 16 here: every stream stays 16-byte aligned)
        if (obs.base_dev) expand_base_u8(st, rb, static_cast<uint8_t*>(obs.base_dev) + (size_t)env0 * st.L, nE, tid, nthr);
        if (obs.color_dev) expand_color_u8(st, rc, static_cast<uint8_t*>(obs.color_dev) + (size_t)env0 * st.S, nE, tid, nthr);
        if (obs.info_dev) expand_info_i32(st, rz, tile_puzzle, static_cast<int32_t*>(obs.info_dev) + (size_t)env0 * st.S, nE, tid, nthr);
        return;
    }
    if (obs.base_dev) expand_base(st, rb, static_cast<int32_t*>(obs.base_dev) + (size_t)env0 * st.L, nE, tid, nthr);
    if (obs.color_dev) expand_color(st, rc, static_cast<int32_t*>(obs.color_dev) + (size_t)env0 * st.S, nE, tid, nthr);
    if (obs.info_dev) expand_info(st, rz, tile_puzzle, static_cast<int64_t*>(obs.info_dev) + (size_t)env0 * st.S, nE, tid, nthr);
}

// Stage 1 for ONE env by one warp (tile = 1: k_post's rewrite of an env that began a new episode): the static words
// of all three streams in one pass over a combined index, so that a lane has all its global loads in flight at once
// (stream by stream, as encode_tile_stage1 does for a CTA, a lane went through up to six dependent L2 round trips:
// delta step 38.2 -> 36.5 us), the visited words and the agent bit OR-ed in on the way.  No tails: nE = 1.
SPARC_HD void encode_env_stage1(const StoreDev& st, const uint32_t* head, uint32_t* flats, const SparcObsOut& obs,
                                int tid, int nthr) {
    const uint32_t pz = head[SPARC_DYN_STRIDE];
    const uint32_t ab = head[8];
    const int Wb = st.W + 2, Wc = st.Wc + 2, Wz = st.nv + 2, R = Wb + Wc + Wz;
    for (int base = tid; base < R; base += 4 * nthr) {
        uint32_t w[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            int k = base + u * nthr;
            w[u] = 0u;
            if (k < Wb) {
                if (k < st.W && obs.base_dev) w[u] = sp_ldg(st.sbits + (pz * (uint32_t)st.W + (uint32_t)k));
            } else if ((k -= Wb) < Wc) {
                if (k < st.Wc && obs.color_dev) w[u] = sp_ldg(st.color4 + (pz * (uint32_t)st.Wc + (uint32_t)k));
            } else if ((k -= Wc) < st.nv && obs.info_dev) {
                w[u] = sp_ldg(st.info_nz + (pz * (uint32_t)st.nv + (uint32_t)k));
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int k = base + u * nthr;
            if (k >= R) continue;
            if (k < st.W) w[u] |= (k < st.nv ? head[k] : 0u) | ((int)(ab >> 5) == k ? 1u << (ab & 31) : 0u);
            flats[k] = w[u];
        }
    }
}

// One env's whole observation from rows staged with tile = 1 (k_encode_list: the few envs per step that began a
// new episode under update = SPARC_OBS_DELTA).  Element-wise stores, consecutive threads -> consecutive elements:
// an env's streams start at 8- or 4-byte boundaries only (L*4 = 6,760 bytes on the 13x13 / 10-plane slab).
template <int FMT>
SPARC_HD void encode_env_elementwise(const StoreDev& st, const uint32_t* head, const uint32_t* flats, long long env,
                                     const SparcObsOut& obs, int tid, int nthr) {
    typedef typename ObsTypes<FMT>::base_t base_t;
    typedef typename ObsTypes<FMT>::info_t info_t;
    const uint32_t puzzle = head[SPARC_DYN_STRIDE];
    const uint32_t* rb = flats;
    const uint32_t* rc = rb + (st.W + 2);
    const uint32_t* rz = rc + (st.Wc + 2);
    if (obs.base_dev) {
        base_t* out = static_cast<base_t*>(obs.base_dev) + (size_t)env * st.L;
        for (int e = tid; e < st.L; e += nthr) out[e] = (base_t)(row_fetch(rb, (uint32_t)e, st.L, st.W + 2, st.inv_L) & 1u);
    }
    if (obs.color_dev) {
        base_t* out = static_cast<base_t*>(obs.color_dev) + (size_t)env * st.S;
        for (int e = tid; e < st.S; e += nthr) out[e] = (base_t)(row_fetch(rc, 4u * (uint32_t)e, 4 * st.S, st.Wc + 2, st.inv_4S) & 15u);
    }
    if (obs.info_dev) {
        info_t* out = static_cast<info_t*>(obs.info_dev) + (size_t)env * st.S;
        for (int e = tid; e < st.S; e += nthr) {
            const uint32_t m = row_fetch(rz, (uint32_t)e, st.S, st.nv + 2, st.inv_S) & 1u;
            out[e] = m ? (info_t)sp_ldg(st.info + (size_t)puzzle * st.S + e) : (info_t)0;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// one launch = reset or step of all envs (+ observation rewrite)
// ---------------------------------------------------------------------------------------------
struct KParams {
    StoreDev st;
    StateDev sd;
    StepCfg cfg;
    int mode;                // 0 = step, 1 = reset
    int tile;                // envs per CTA
    const void* actions;
    int act_dtype;
    const int32_t* reset_idx;
    const uint8_t* reset_mask;
    SparcStepOut out;
    SparcObsOut obs;
    uint32_t* vq_count;      // deferred terminal validation: number of queued envs (nullptr = validate inline)
    uint32_t* vq;            // [n_envs][SPARC_VQ_WORDS]: env, puzzle, x | y << 8, pad, visited[8]
    uint32_t* hq_count;      // hard items: snapshots whose exact-fit search outgrew the validator's node budget (k_val_hard)
#if defined(__CUDACC__)
    uint4* hq;
#else
    void* hq;
#endif
    int fit_budget;          // nodes an inline exact-fit search may visit before the snapshot becomes a hard item (0 = no limit)
    uint32_t* rq_count;      // delta observations: number of envs that began a new episode this step (nullptr in the
    uint32_t* rq;            // fused kernel, which decides per tile); [n_envs] their indices -> k_encode_list
    // sparc_step_host on the fused kernel: the host waits for done_flag (mapped host memory) to become done_seq instead
    // of synchronising the stream; every CTA counts itself in done_count once its per-env results are written
    uint32_t* done_flag;
    uint32_t* done_count;
    uint32_t done_seq;
    uint8_t* packed_out;     // per env: reward code | terminated << 3 | truncated << 4 (device memory); the last CTA to
    float* host_reward;      // finish its dynamics expands them into these arrays in mapped host memory (n rounded up
    uint8_t* host_term;      // to 16 entries each)
    uint8_t* host_trunc;
};

SPARC_HD long long load_action(const void* p, int dtype, long long i) {
    if (dtype == SPARC_ACT_U8) return reinterpret_cast<const uint8_t*>(p)[i];
    if (dtype == SPARC_ACT_I32) return reinterpret_cast<const int32_t*>(p)[i];
    return reinterpret_cast<const long long*>(p)[i];
}

SPARC_HD float reward_value(int code) {
    return code == SPARC_R_ON_PATH ? 0.01f : code == SPARC_R_OFF_PATH ? -0.01f
         : code == SPARC_R_WIN ? 1.0f : code == SPARC_R_LOSE ? -1.0f : 0.0f;
}

// Phase A for env i (one thread): reset / autoreset / frozen / live step, terminal validation,
// per-env outputs.  `scratch` = 9 words the validator may use.  Returns true when `stats`
// (SPARC_N_STATS counters) were filled.  On exit `e` is the state the observation must show.
// In-place observation update of one env after a move (update = SPARC_OBS_DELTA; SG.py:1146-1188 mutates the aliased
// planes the same way): agent_location leaves (ox, oy) and enters (x, y); visited is set at (x, y) and, at (ox, oy),
// keeps its value after a fresh move or is cleared by a traceback undo -- i.e. it becomes the state's bit.
template <typename T>
SPARC_HD void obs_delta_planes(T* base, const StoreDev& st, long long i, int ox, int oy, const EnvRegs& e) {
    T* b = base + (size_t)i * st.L;
    const int o = ox * st.SY + oy, c = e.x * st.SY + e.y;
    b[o] = (T)board_bit(e.v, o);                 // plane 0: visited
    b[c] = (T)1;
    b[2 * st.S + o] = (T)0;                      // plane 2: agent_location
    b[2 * st.S + c] = (T)1;
}

// `full` (optional): set when the env began a new episode in this call, i.e. its whole observation changed.
// MAY_DELTA = false compiles the in-place observation update out (k_dyn of the full-rewrite step: 62 registers
// instead of 66, i.e. 8 instead of 7 resident CTAs per SM).
enum { PA_NEW_EPISODE = 1, PA_SNAPSHOT = 2 };
#define SPARC_NO_ARG (-0x7fffffffffffffffll - 1)
template <bool INLINE_VALIDATION = true, bool MAY_DELTA = true>
SPARC_HD bool phase_a(const KParams& p, long long i, EnvRegs& e, uint32_t* scratch, unsigned long long* stats,
                      int* tile_flags = nullptr, uint32_t* snap = nullptr, long long arg = SPARC_NO_ARG,
                      StepResult* result = nullptr) {
    // arg: this env's action (step) or puzzle index (reset) when it came with the launch instead of through
    // p.actions / p.reset_idx (the host-buffer calls on launch-bound batches).  result: the step's outcome for a
    // caller that writes the per-env outputs itself.
    // tile_flags (the fused k_env): PA_NEW_EPISODE | PA_SNAPSHOT of this env.  snap: SPARC_VQ_WORDS words of the CTA's
    // shared memory; a state that needs validating is stored there instead of being validated by this thread, and
    // another warp of the CTA validates it beside the encoder (validate_snapshot).
    const StoreDev& st = p.st;
    const long long n = p.cfg.n_envs;
    load_env(e, p.sd, st, n, i);
    // (Prefetching the header and the walk board of the puzzle a same-step autoreset would move to -- the two extra
    // dependent L2 round trips of the 1-2 % of envs that end an episode -- changed nothing measurable: config 2 8.32 vs
    // 8.36 us, delta 35.5-36.6 either way.)
    StepResult r;
    r.reward_code = SPARC_R_ZERO; r.terminated = 0; r.truncated = 0; r.legal_mask = 0; r.rule_flags = 0;
    bool need_legal = false, have_stats = false, new_episode = false, snapped = false;
    const int ox = (int)e.x, oy = (int)e.y;
    if (p.mode == 1) {                                                   // ---- reset (SG.py:1057-1108)
        const bool sel = p.reset_mask ? p.reset_mask[i] != 0 : true;
        if (sel) {
            // indices are taken modulo the pool, Python style (-1 = the last puzzle); host wrappers range-check first
            int pzi = arg != SPARC_NO_ARG ? (int)(arg % st.n_puzzles) : p.reset_idx ? p.reset_idx[i] % st.n_puzzles : (int)e.puzzle;
            if (pzi < 0) pzi += st.n_puzzles;
            const uint32_t pz = (uint32_t)pzi;
            env_reset(e, st, pz);
            store_env(e, p.sd, st, n, i);
            new_episode = true;
        }
        need_legal = true;
        r.terminated = (e.flags & FL_TERM) != 0; r.truncated = (e.flags & FL_TRUNC) != 0;
    } else if (e.flags & FL_DONE) {                                      // ---- finished episode
        if (p.cfg.autoreset == SPARC_AUTORESET_NEXT_STEP) {
            env_reset(e, st, (e.puzzle + (uint32_t)p.cfg.stride) % (uint32_t)st.n_puzzles);
            store_env(e, p.sd, st, n, i);
            new_episode = true;
        } else {                                                         // frozen until sparc_reset
            r.terminated = (e.flags & FL_TERM) != 0; r.truncated = (e.flags & FL_TRUNC) != 0;
        }
        need_legal = true;
    } else {                                                             // ---- live step
        const long long a = arg != SPARC_NO_ARG ? arg : load_action(p.actions, p.act_dtype, i);
        env_step(e, st, p.sd, p.cfg, i, a, r);
        const bool done = r.terminated || r.truncated;
        if (p.cfg.validate_mode == SPARC_VALIDATE_EVERY_STEP ||
            (p.cfg.validate_mode == SPARC_VALIDATE_TERMINAL && done)) {
            if (snap) {
                snap[0] = (uint32_t)i; snap[1] = e.puzzle; snap[2] = (uint32_t)e.x | ((uint32_t)e.y << 8); snap[3] = done;     // [3]: counts for RULES_PASS
#pragma unroll
                for (int k = 0; k < 8; k++) snap[4 + k] = e.v[k];
                snapped = true;
            } else if (!INLINE_VALIDATION || (p.vq_count && p.cfg.validate_mode == SPARC_VALIDATE_TERMINAL)) {
                // Terminal states are rare (1-2 % of the envs per step) and validating one inline idles
                // the 31 other lanes of its warp for the whole flood fill: queue a snapshot instead;
                // validate_queued() runs densely packed, beside the observation encoder.
#if defined(__CUDA_ARCH__)
                const uint32_t slot = atomicAdd(p.vq_count, 1u);
#else
                const uint32_t slot = (*p.vq_count)++;
#endif
                uint32_t* q = p.vq + (size_t)slot * SPARC_VQ_WORDS;
                q[0] = (uint32_t)i; q[1] = e.puzzle; q[2] = (uint32_t)e.x | ((uint32_t)e.y << 8); q[3] = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) q[4 + k] = e.v[k];
            } else if (INLINE_VALIDATION) {
#pragma unroll
                for (int k = 0; k < 8; k++) scratch[k] = e.v[k];
                scratch[8] = 0;
                r.rule_flags = validate_env(st, e.puzzle, scratch, e.x, e.y, nullptr) | SPARC_RULE_VALID;
            }
        }
#pragma unroll
        for (int k = 0; k < SPARC_N_STATS; k++) stats[k] = 0;
        have_stats = true;
        stats[SPARC_STAT_STEPS] = 1;
        if (done) {
            stats[SPARC_STAT_EPISODES] = 1;
            stats[SPARC_STAT_SOLVED] = r.reward_code == SPARC_R_WIN;
            stats[SPARC_STAT_FAILED] = r.reward_code == SPARC_R_LOSE;
            stats[SPARC_STAT_TRUNCATED] = r.truncated;
            stats[SPARC_STAT_RETURN_CENTS] = (unsigned long long)(long long)e.ep_ret;
            stats[SPARC_STAT_LENGTH] = e.step;
            stats[SPARC_STAT_RULES_PASS] = (r.rule_flags & SPARC_RULE_ALL) != 0;
            if (p.cfg.autoreset == SPARC_AUTORESET_SAME_STEP) {
                env_reset(e, st, (e.puzzle + (uint32_t)p.cfg.stride) % (uint32_t)st.n_puzzles);
                need_legal = true;
                new_episode = true;
            } else if (p.cfg.autoreset == SPARC_AUTORESET_CONTINUE) {
                // the reference has no guard (Q4): the episode goes on, the next step is a live step again
            } else {
                e.flags = FL_DONE | (r.terminated ? FL_TERM : 0) | (r.truncated ? FL_TRUNC : 0);
            }
        }
        store_env(e, p.sd, st, n, i);
    }
    if (need_legal) r.legal_mask = env_legal_now(e, st, p.sd, p.cfg, i);
    if (tile_flags) *tile_flags = (new_episode ? PA_NEW_EPISODE : 0) | (snapped ? PA_SNAPSHOT : 0);
    if (MAY_DELTA && p.mode == 0 && p.obs.update == SPARC_OBS_DELTA) {   // persistent observation, touched in place
        if (new_episode) {
            if (p.rq_count) {                                            // whole observation: k_encode_list rewrites it
#if defined(__CUDA_ARCH__)
                const uint32_t slot = atomicAdd(p.rq_count, 1u);
#else
                const uint32_t slot = (*p.rq_count)++;
#endif
                p.rq[slot] = (uint32_t)i;
            }
        } else if (p.obs.base_dev && ((int)e.x != ox || (int)e.y != oy)) {
            if (p.obs.format == SPARC_OBS_COMPACT) obs_delta_planes(static_cast<uint8_t*>(p.obs.base_dev), st, i, ox, oy, e);
            else obs_delta_planes(static_cast<int32_t*>(p.obs.base_dev), st, i, ox, oy, e);
        }
    }
    if (result) *result = r;
    // per-env outputs (SG.py:1238)
    if (p.packed_out) p.packed_out[i] = (uint8_t)(r.reward_code | (r.terminated << 3) | (r.truncated << 4));
    if (p.out.reward_dev) p.out.reward_dev[i] = reward_value(r.reward_code);
    if (p.out.reward_code_dev) p.out.reward_code_dev[i] = (int8_t)r.reward_code;
    if (p.out.terminated_dev) p.out.terminated_dev[i] = (uint8_t)r.terminated;
    if (p.out.truncated_dev) p.out.truncated_dev[i] = (uint8_t)r.truncated;
    if (p.out.legal_mask_dev) p.out.legal_mask_dev[i] = (uint8_t)r.legal_mask;
    if (p.out.rule_flags_dev && !snapped) p.out.rule_flags_dev[i] = (uint16_t)r.rule_flags;     // (snapped: validate_snapshot writes them)
    if (p.out.step_count_dev) p.out.step_count_dev[i] = (int32_t)e.step;
    if (p.out.agent_xy_dev) { p.out.agent_xy_dev[2 * i] = (uint8_t)e.x; p.out.agent_xy_dev[2 * i + 1] = (uint8_t)e.y; }
    return have_stats;
}

// deferred terminal validation of queue entry `slot` (one thread): writes the env's rule flags, returns 1 when all
// rules are satisfied (for the RULES_PASS statistic), 0 when not, -1 when an exact-fit search outgrew `fit_budget`
// (nothing written: the snapshot becomes a hard item)
SPARC_HD int validate_snapshot(const KParams& p, const uint32_t* q, int fit_budget = 0) {
    uint32_t vd[SPARC_DYN_STRIDE];
#pragma unroll
    for (int k = 0; k < 8; k++) vd[k] = q[4 + k];
    vd[8] = 0;
    int f = validate_env(p.st, q[1], vd, (int)(q[2] & 0xff), (int)((q[2] >> 8) & 0xff), nullptr, nullptr, fit_budget);
    if (f < 0) return -1;
    f |= SPARC_RULE_VALID;
    if (p.out.rule_flags_dev) p.out.rule_flags_dev[q[0]] = (uint16_t)f;
    return (f & SPARC_RULE_ALL) != 0;
}
SPARC_HD int validate_queued(const KParams& p, uint32_t slot, int fit_budget = 0) {
    return validate_snapshot(p, p.vq + (size_t)slot * SPARC_VQ_WORDS, fit_budget);
}
