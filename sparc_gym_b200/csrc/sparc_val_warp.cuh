// sparc_val_warp.cuh -- warp-cooperative rule validation of ONE queued terminal state (k_val).  Device only.
//
// Same verdicts as validate_env() in sparc_core.cuh (SG.py:374-950; the scalar version stays the one the fused
// small-batch kernel, sparc_validate and the CPU build of the tests run), laid out for 32 lanes:
//   * the puzzle's validator record (SPARC_VREC_*, built by sparc_store_create) is ONE level of loads after the
//     queue entry: lane l reads row (l & 7) = word (l & 7) of the nine 16-stride node boards, the header and the cell
//     boards the region rules need, colour (l & 7) -- six independent 16-byte loads per lane, 768 bytes per puzzle;
//   * node boards live one word per lane (lanes 8..31 replicate lanes 0..7, so no lane is predicated off and every
//     shuffle / vote is warp-wide): a flood sweep is one in-lane row fill, two shuffles and a vote; gaps / dots /
//     triangles are one word each and one redux;
//   * the per-region rules run one colour per lane (squares, stars: one ballot) and one shape slot per lane (instance
//     counts and the area balance: one ballot and one redux.add);
//   * the exact polyomino fit runs 32 copies of the depth-first search on disjoint shares of the tree (fit_step with
//     code = lane, see sparc_core.cuh), reconverging after every node; the first copy that finds a fit ends all.
#pragma once
#include "sparc_core.cuh"

// validator record: [puzzle][row k = 0..7][24 words]
//   0 GAPS16[k]  1 CELL16[k]  2 VALID16[k]  3 DOT16[k]  4..7 TRI1..4[k]  8 TRIBAD[k]
//   9 X | Y << 8 | start_x << 16 | start_y << 24   10 target_x | target_y << 8 | n_shapes << 16 | flags << 24   11 shape_off
//   12,13 STAR  14,15 SQUARE  16,17 SYM1  18,19 SYM2  20,21 COLOR(k+1)  22,23 OR of the eight colour boards
#define SPARC_VREC_ROW_WORDS 24
#define SPARC_VREC_WORDS (8 * SPARC_VREC_ROW_WORDS)

#if defined(__CUDACC__)

#define VW_FULL 0xffffffffu

__device__ __forceinline__ uint64_t vw_u64(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

// 32 copies of the search, copy `lane` on its share of the tree; warp-uniform result
__device__ __forceinline__ bool exact_fit_warp(uint64_t region, int w, int h, const SparcShape* shapes, int ns,
                                               const Cnt8& n_poly, const Cnt8& n_ylop, int lane) {
    FitState f;
    int r = fit_init(f, region, w, h, shapes, ns, n_poly, n_ylop);            // the same on every lane
    if (r >= 0) return r != 0;
    bool live = true, found = false;
    for (;;) {
        if (live) {
            r = fit_step(f, shapes, ns, lane, 5);
            if (r >= 0) { live = false; found = r != 0; }
        }
        const unsigned v = __reduce_or_sync(VW_FULL, (found ? 2u : 0u) | (live ? 1u : 0u));   // also reconverges the copies
        if (v & 2u) return true;
        if (!(v & 1u)) return false;
    }
}

#if defined(VAL_PROFILE)
#define VW_CLOCK() clock64()
#else
#define VW_CLOCK() 0ll
#endif

// qw: lane l < 12 holds word l of the queue entry (env, puzzle, x | y << 8, pad, visited[8] dense), the others 0.
// Returns the SPARC_RULE_* flags (without SPARC_RULE_VALID), the same on every lane.
__device__ __forceinline__ int validate_warp(const StoreDev& st, uint32_t qw, int lane) {
    const long long vp_t0 = VW_CLOCK();
    long long vp_pro = 0, vp_flood = 0, vp_fit = 0, vp_tmp = 0;
    (void)vp_t0; (void)vp_pro; (void)vp_flood; (void)vp_fit; (void)vp_tmp;
    const int k = lane & 7;
    const uint32_t puzzle = __shfl_sync(VW_FULL, qw, 1);
    const uint32_t axy = __shfl_sync(VW_FULL, qw, 2);
    const uint4* row = reinterpret_cast<const uint4*>(st.vrec + ((size_t)puzzle * 8 + k) * SPARC_VREC_ROW_WORDS);
    const uint4 r0 = __ldg(row), r1 = __ldg(row + 1), r2 = __ldg(row + 2);
    const uint4 r3 = __ldg(row + 3), r4 = __ldg(row + 4), r5 = __ldg(row + 5);
    // visited: dense slab order (queue words 4..11) -> word k of the 16-stride board (rows x = 2k, 2k + 1)
    uint32_t V = 0;
    const uint32_t rowmask = (1u << st.SY) - 1u;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const int x = 2 * k + half, pos = x * st.SY, idx = pos >> 5;
        const uint32_t lo = __shfl_sync(VW_FULL, qw, 4 + idx), hi = __shfl_sync(VW_FULL, qw, 5 + idx);   // lane 12 holds 0
        const uint32_t bits = __funnelshift_r(lo, hi, pos & 31) & rowmask;
        V |= (x < st.SX ? bits : 0u) << (16 * half);
    }
    const uint32_t G = r0.x, C = r0.y, VAL = r0.z, DOT = r0.w;
    const uint32_t h0 = r2.y, h1 = r2.z, shape_off = r2.w;
    int flags = SPARC_RULE_PATH_NOT_CROSSING;      // path nodes are distinct by construction (SG.py:498)
    if ((axy & 0xffu) == (h1 & 0xffu) && ((axy >> 8) & 0xffu) == ((h1 >> 8) & 0xffu)) flags |= SPARC_RULE_REACHED_TARGET;   // SG.py:488
    {   // gaps (SG.py:508), dots (SG.py:521), triangles (SG.py:617-646): one word per lane, one redux
        const uint32_t Vm = __shfl_up_sync(VW_FULL, V, 1, 8), Vp = __shfl_down_sync(VW_FULL, V, 1, 8);
        const uint32_t a = (V << 16) | (k > 0 ? Vm >> 16 : 0u);               // node (x-1, y)
        const uint32_t b = (V >> 16) | (k < 7 ? Vp << 16 : 0u);               // node (x+1, y)
        const uint32_t c = (V << 1) & 0xFFFEFFFEu;                            // node (x, y-1)
        const uint32_t d = (V >> 1) & 0x7FFF7FFFu;                            // node (x, y+1)
        const uint32_t s1 = a ^ b, c1 = a & b, s2 = c ^ d, c2 = c & d;
        const uint32_t b0 = s1 ^ s2, cr = s1 & s2, b1 = c1 ^ c2 ^ cr, b2 = c1 & c2;
        const uint32_t eq1 = b0 & ~b1 & ~b2, eq2 = ~b0 & b1 & ~b2, eq3 = b0 & b1, eq4 = b2;
        const uint32_t bad = (r1.x & ~eq1) | (r1.y & ~eq2) | (r1.z & ~eq3) | (r1.w & ~eq4) | r2.x;
        const unsigned viol = __reduce_or_sync(VW_FULL, ((G & V) ? 1u : 0u) | ((DOT & ~V) ? 2u : 0u) | (bad ? 4u : 0u));
        if (!(viol & 1u)) flags |= SPARC_RULE_NO_GAP_VIOLATIONS;
        if (!(viol & 2u)) flags |= SPARC_RULE_ALL_DOTS;
        if (!(viol & 4u)) flags |= SPARC_RULE_TRIANGLES;
    }
    const int hflags = (int)(h1 >> 24), n_shapes = (int)((h1 >> 16) & 0xffu);
    bool sq_ok = true, star_ok = true, poly_ok = true;
    vp_pro = VW_CLOCK() - vp_t0;
    if (hflags & 2) {                      // some region rule can fail: segment (SG.py:422-454; passable per Q6)
        const uint64_t STAR = vw_u64(r3.x, r3.y), SQ = vw_u64(r3.z, r3.w);
        const uint64_t SB0 = vw_u64(r4.x, r4.y), SB1 = vw_u64(r4.z, r4.w);
        const uint64_t COLk = vw_u64(r5.x, r5.y), ANYCOL = vw_u64(r5.z, r5.w);
        const SparcShape* shapes = st.shapes + shape_off;
        uint64_t sh_anchors = 0, sh_poly = 0, sh_ylop = 0;                    // slot `lane`
        if (lane < n_shapes) {
            sh_anchors = __ldg(&shapes[lane].anchors);
            const ulonglong2 pc = __ldg(reinterpret_cast<const ulonglong2*>(&shapes[lane].poly_cells));
            sh_poly = pc.x; sh_ylop = pc.y;
        }
        const uint32_t Pm = VAL & (C | ~(G | V));
        uint32_t unl = C;
        const int w = (int)((h0 & 0xffu) - 1u) >> 1, hh = (int)(((h0 >> 8) & 0xffu) - 1u) >> 1;
        for (;;) {
            const unsigned nz = __ballot_sync(VW_FULL, unl != 0u) & 0xffu;
            if (!nz) break;
            const int k0 = __ffs((int)nz) - 1;
            uint32_t f = (k == k0) ? (unl & (0u - unl)) : 0u;                  // lowest unlabelled cell: raster order
            vp_tmp = VW_CLOCK();
            for (;;) {                                                         // flood to fix-point, two sweeps per vote
                bool ch = false;
#pragma unroll
                for (int rep = 0; rep < 2; rep++) {
                    const uint32_t g = row_fill(Pm, f);
                    const uint32_t gm = __shfl_up_sync(VW_FULL, g, 1, 8), gp = __shfl_down_sync(VW_FULL, g, 1, 8);
                    const uint32_t up = (g << 16) | (k > 0 ? gm >> 16 : 0u);
                    const uint32_t dn = (g >> 16) | (k < 7 ? gp << 16 : 0u);
                    const uint32_t nw = (g | up | dn) & Pm;
                    ch = nw != f;                                              // (a change in sweep 1 implies one in sweep 2 or a fix-point there)
                    f = nw;
                }
                if (!__any_sync(VW_FULL, ch)) break;
            }
            vp_flood += VW_CLOCK() - vp_tmp;
            const uint32_t cells = f & C;
            unl &= ~cells;
            // component -> rule-cell board (bit 8cx + cy): lane k holds column cx = k in the high half of its word
            uint32_t t = (cells >> 17) & 0x5555u;
            t = (t | (t >> 1)) & 0x3333u; t = (t | (t >> 2)) & 0x0F0Fu; t = (t | (t >> 4)) & 0x00FFu;
            const uint32_t Rlo = __reduce_or_sync(VW_FULL, k < 4 ? t << (8 * k) : 0u);
            const uint32_t Rhi = __reduce_or_sync(VW_FULL, (k >= 4 && k < 7) ? t << (8 * (k - 4)) : 0u);
            const uint64_t R = vw_u64(Rlo, Rhi);
            // squares (SG.py:534-552): at most one distinct non-zero colour per region; stars (SG.py:554-614): every
            // star colour totals exactly 2 symbols in the region, no colourless star (SG.py:580-582).  Colour k per lane.
            const uint64_t rc = R & COLk;
            const bool sq_here = (SQ & rc) != 0ull;
            const bool star_bad = (STAR & rc) != 0ull && (__popcll(SB0 & rc) + 2 * __popcll(SB1 & rc) != 2);
            const unsigned vote = __ballot_sync(VW_FULL, lane < 8 ? sq_here : star_bad);
            if (__popc(vote & 0xffu) > 1) sq_ok = false;
            if ((vote >> 8) || (STAR & R & ~ANYCOL)) star_ok = false;
            // polys / ylops (SG.py:648-710): slot `lane`
            if (n_shapes && poly_ok) {
                const int pa = __popcll(sh_poly & R), yb = __popcll(sh_ylop & R);
                const unsigned inst = __ballot_sync(VW_FULL, (pa | yb) != 0);
                if (inst) {
                    const int net = __reduce_add_sync(VW_FULL, (pa - yb) * (int)(sh_anchors >> 56));
                    if (net != __popcll(R)) poly_ok = false;                   // SG.py:679-682
                    else {
                        Cnt8 np = {{0, 0, 0, 0}}, nyl = {{0, 0, 0, 0}};
                        for (unsigned m = inst; m; m &= m - 1) {
                            const int s = __ffs((int)m) - 1;
                            cnt8_add(np, s, __shfl_sync(VW_FULL, pa, s)); cnt8_add(nyl, s, __shfl_sync(VW_FULL, yb, s));
                        }
                        vp_tmp = VW_CLOCK();
                        if (!exact_fit_warp(R, w, hh, shapes, n_shapes, np, nyl, lane)) poly_ok = false;
                        vp_fit += VW_CLOCK() - vp_tmp;
                    }
                }
            }
        }
    }
    if (sq_ok) flags |= SPARC_RULE_SQUARES;
    if (star_ok) flags |= SPARC_RULE_STARS;
    if (poly_ok) flags |= SPARC_RULE_POLY_YLOP;
    if ((flags & 0xFF) == 0xFF) flags |= SPARC_RULE_ALL;                       // SG.py:932-935
#if defined(VAL_PROFILE)
    if (lane == 0) {
        const unsigned long long tot = (unsigned long long)(clock64() - vp_t0);
        atomicAdd(&g_valprof[0], tot); atomicAdd(&g_valprof[1], (unsigned long long)vp_pro);
        atomicAdd(&g_valprof[2], (unsigned long long)vp_flood); atomicAdd(&g_valprof[3], (unsigned long long)vp_fit);
        atomicAdd(&g_valprof[4], 1ull); atomicMax(&g_valprof[5], tot);
        atomicMax(&g_valprof[6], (unsigned long long)vp_flood); atomicMax(&g_valprof[7], (unsigned long long)vp_fit);
    }
#endif
    return flags;
}

#endif  // __CUDACC__
