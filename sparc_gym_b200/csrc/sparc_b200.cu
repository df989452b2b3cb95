// sparc_b200.cu -- sm_100a kernels + the C ABI of include/sparc_b200.h.
//
// One env step of a big batch (>= 8,192 envs) is three launches:
//   k_dyn     one env per thread: legality, move / traceback undo, trie-cursor reward, truncation,
//             statistics, autoreset; terminal states are queued as 48-byte snapshots
//   k_val     walks that queue (rule validation: flood fill, colour rules, exact polyomino fit) on a side
//             stream, beside ...
//   k_encode  state -> observation: a CTA owns a tile of 8 consecutive envs, stages their bit-packed planes
//             in shared memory and expands them into the caller's int32/int64 tensors with coalesced
//             16-byte streaming stores.  HBM-write bound: X*Y*(4P+4+8) bytes per env-step.
// With update = SPARC_OBS_DELTA the step is k_dyn (touches the changed cells in place) + k_post (warp-cooperative
// validation of the queued terminal states and the full rewrite of the envs that began a new episode).
// Small batches use the fused k_env (one launch, one tile per CTA: phase A in warp 0, then warps 0-2 encode while warp 3
// validates the tile's terminal states); k_env_host is the same launch for sparc_step_host (actions in the kernel
// parameters, results + completion flag into mapped host memory); k_info is the single-env drop-in's whole step.
// The device functions themselves (dynamics, validator, encoder) live in sparc_core.cuh, the warp-cooperative
// validator in sparc_val_warp.cuh.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>          // header-only NVTX 3: ranges around the ABI calls when SPARC_NVTX=1 (Nsight Systems / ncu --nvtx)
#include <stdio.h>
#include <string.h>
#include <new>

#include "sparc_core.cuh"
#include "sparc_val_warp.cuh"

#define CTA_THREADS 128
#ifndef SPARC_MINBLOCKS          // resident CTAs per SM the register allocator must allow.  8 -> 64
#define SPARC_MINBLOCKS 8        // registers: the rarely-run validator spills, the streaming encoder
#endif                           // gets 2x the resident warps (measured: 0.322 -> 0.219 ms/step, config 4)

static thread_local char g_err[512] = "";
struct NvtxRange {           // off unless SPARC_NVTX=1, so the launch path pays one predictable branch
    bool on;
    explicit NvtxRange(const char* name) {
        static const bool enabled = getenv("SPARC_NVTX") && atoi(getenv("SPARC_NVTX")) != 0;
        on = enabled;
        if (on) nvtxRangePushA(name);
    }
    ~NvtxRange() { if (on) nvtxRangePop(); }
};
static int fail(int code, const char* fmt, const char* detail = "") {
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}
#define CU(call)                                                                         \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) return fail(SPARC_E_CUDA, #call ": %s", cudaGetErrorString(e_)); \
    } while (0)

struct SparcStore {
    int device;
    int max_instances;       // most poly / ylop instances in one puzzle of the pool
    StoreDev d;
    void* blocks[10];
    int plane_ids[SPARC_MAX_PLANES];
};

// sparc_step_host on launch-bound batches: the actions ride in the launch itself, two bits each (<= 8,192 envs = 2 KB of
// kernel parameters) -- no PCIe read in front of the dynamics; see k_env_host below
#define SPARC_PACKED_MAX_ENVS 8192
struct PackedActions { uint32_t w[SPARC_PACKED_MAX_ENVS / 16]; };

struct SparcEnv {
    SparcStore* store;
    StepCfg cfg;
    StateDev s;
    int tile;                // envs per CTA
    int split_chunks;        // >0: k_dyn / k_encode pipeline over this many env chunks
    int enc_tile;            // envs per k_encode CTA
    cudaStream_t side;       // high-priority stream of k_dyn (multi-chunk) / k_val
    cudaEvent_t ev_start, ev_dyn[16], ev_val;
    cudaStream_t copy_stream;   // sparc_step_host: result copies run beside k_encode
    cudaEvent_t ev_dyn_done, ev_copy, ev_act;
    struct { float* reward; uint8_t* term; uint8_t* trunc; bool active; } hostcopy;
    uint32_t* vq_count;      // deferred-validation queue (split path, validate = terminal)
    uint32_t* vq;
    uint32_t* hq_count;      // hard items (exact-fit searches over budget): items, shares handed out, finished CTAs
    uint4* hq;               // [n_envs]
    int hard_search;         // the pool can produce big exact-fit searches: node budgets + k_val_hard (see push_hard_item)
    uint32_t* rq_count;      // update = SPARC_OBS_DELTA: envs that began a new episode this step (count, finished CTAs)
    uint32_t* rq;            // [n_envs]
    int val_grid;            // CTAs of k_val beside the full-rewrite encoder
    int val_grid_delta;      // most CTAs of k_post (update = SPARC_OBS_DELTA)
    int64_t launches;
    char* info_stage;        // device copy of the record of sparc_step_info_host (allocated on first use)
    int32_t* int_stage;      // its int32 actions / puzzle indices
    uint8_t* act_stage;      // device staging for the *_host call
    float* rew_stage;
    uint8_t* term_stage;
    uint8_t* trunc_stage;
    // sparc_step_host, fused kernel: completion flag in mapped host memory (see KParams::done_flag) and the device
    // addresses of the caller's pinned buffers as of the last call
    volatile uint32_t* done_host;
    uint32_t* done_dev;
    uint32_t* done_count;
    uint32_t done_seq;
    bool notify;             // this launch is k_env_host: packed actions in, packed results + done flag out
    uint8_t* packed_dev;     // [n_envs rounded up to 16] per-env result bytes
    char* info_host;         // mapped host memory: the record of sparc_*_info_host (k_info; allocated on first use)
    char* info_host_dev;
    char* res_host;          // mapped host memory: reward f32[nb] | terminated u8[nb] | truncated u8[nb], nb = n_envs rounded up to 16
    char* res_host_dev;
    size_t res_nb;
    PackedActions pa;
};

// dynamic shared memory of both kernels: tile heads (dyn rows + puzzle ids) | flat strings
extern __shared__ __align__(128) uint32_t smem[];

// Episode statistics: SPARC_STAT_ROWS rows of SPARC_N_STATS counters, a CTA adds into row blockIdx mod ROWS and
// k_read_stats sums the rows.  (One row for everybody meant ~12,000 atomics per step on eight addresses; they
// serialise in L2 and k_dyn is latency-bound as it is.)
#ifndef SPARC_STAT_ROWS
#define SPARC_STAT_ROWS 1024
#endif
__device__ __forceinline__ void reduce_stats(const KParams& p, bool have, const unsigned long long* my) {
    // steps / episodes / solved / failed / truncated / rules_pass are 0-or-1 per env: ballots, not shuffle sums;
    // only a warp in which an episode ended (40 % of them at 1.6 % per env) reduces anything beyond `steps`
    const unsigned full = 0xffffffffu;
    const bool lead = (threadIdx.x & 31) == 0;
    unsigned long long* row = p.sd.stats + (size_t)(blockIdx.x % SPARC_STAT_ROWS) * SPARC_N_STATS;
    const unsigned live = __ballot_sync(full, have && my[SPARC_STAT_STEPS]);
    if (lead && live) atomicAdd(row + SPARC_STAT_STEPS, (unsigned long long)__popc(live));
    const bool ended = have && my[SPARC_STAT_EPISODES];
    const unsigned done = __ballot_sync(full, ended);
    if (!done) return;                                                          // warp-uniform
    const unsigned solved = __ballot_sync(full, ended && my[SPARC_STAT_SOLVED]);
    const unsigned failed = __ballot_sync(full, ended && my[SPARC_STAT_FAILED]);
    const unsigned trunc = __ballot_sync(full, ended && my[SPARC_STAT_TRUNCATED]);
    const unsigned rules = __ballot_sync(full, ended && my[SPARC_STAT_RULES_PASS]);
    long long ret = ended ? (long long)my[SPARC_STAT_RETURN_CENTS] : 0ll;      // signed (two's complement in the counter)
    unsigned long long len = ended ? my[SPARC_STAT_LENGTH] : 0ull;
#pragma unroll
    for (int off = 16; off; off >>= 1) { ret += __shfl_down_sync(full, ret, off); len += __shfl_down_sync(full, len, off); }
    if (lead) {
        atomicAdd(row + SPARC_STAT_EPISODES, (unsigned long long)__popc(done));
        if (solved) atomicAdd(row + SPARC_STAT_SOLVED, (unsigned long long)__popc(solved));
        if (failed) atomicAdd(row + SPARC_STAT_FAILED, (unsigned long long)__popc(failed));
        if (trunc) atomicAdd(row + SPARC_STAT_TRUNCATED, (unsigned long long)__popc(trunc));
        if (rules) atomicAdd(row + SPARC_STAT_RULES_PASS, (unsigned long long)__popc(rules));
        if (ret) atomicAdd(row + SPARC_STAT_RETURN_CENTS, (unsigned long long)ret);
        atomicAdd(row + SPARC_STAT_LENGTH, len);
    }
}

// -DENV_PROFILE (tools/envprof.py): where one CTA of the fused k_env spends its time.  Thread 0 of every CTA reads
// clock64 after phase A (its own env), after the barrier behind it (= the CTA's slowest env), after stages 1a / 1b and at
// the end of the CTA (behind one more barrier), plus globaltimer at both ends.  g_envprof: 0 CTAs, 1-5 sums of the five
// marks in cycles since the CTA began, 6-10 their maxima, 11 / 12 earliest begin / latest end (ns), 13 sum of begins (low
// 32 bits, ns).  Not compiled otherwise.
#if defined(ENV_PROFILE)
__device__ unsigned long long g_envprof[16];
__device__ __forceinline__ unsigned long long ep_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define EP_BEGIN() const unsigned long long ep_g0 = ep_ns(); const long long ep_t0 = clock64(); long long ep_m[4] = {0, 0, 0, 0}
#define EP_MARK(k) (ep_m[k] = clock64() - ep_t0)
#define EP_END() do { __syncthreads(); if (threadIdx.x == 0) { const unsigned long long ep_e = (unsigned long long)(clock64() - ep_t0); \
    atomicAdd(&g_envprof[0], 1ull); \
    for (int k = 0; k < 4; k++) { atomicAdd(&g_envprof[1 + k], (unsigned long long)ep_m[k]); atomicMax(&g_envprof[6 + k], (unsigned long long)ep_m[k]); } \
    atomicAdd(&g_envprof[5], ep_e); atomicMax(&g_envprof[10], ep_e); \
    atomicMin(&g_envprof[11], ep_g0); atomicMax(&g_envprof[12], ep_ns()); atomicAdd(&g_envprof[13], ep_g0 & 0xffffffffull); } } while (0)
#else
#define EP_BEGIN() ((void)0)
#define EP_MARK(k) ((void)0)
#define EP_END() ((void)0)
#endif

// ---- simple kernel: one CTA = one tile (small batches: every CTA is resident at once) ----------
template <int FMT>
__device__ __forceinline__ void k_env_body(const KParams& p, const uint32_t* packed_actions) {
    const StoreDev& st = p.st;
    const int tid = threadIdx.x;
    const long long env0 = (long long)blockIdx.x * p.tile;
    const long long n = p.cfg.n_envs;
    const int nE = (int)min((long long)p.tile, n - env0);
    uint32_t* head = smem;
    uint32_t* flats = smem + tile_head_words(p.tile);

    uint32_t* tflags = flats + tile_flat_words(st, p.tile);      // [tile] PA_* of every env | [tile][SPARC_VQ_WORDS] snapshots
    uint32_t* snaps = tflags + p.tile;
    EP_BEGIN();
    EnvRegs e;
    unsigned long long my_stats[SPARC_N_STATS];
    bool have_stats = false;
    // (The warps that wait for phase A prefetching the static rows of the env's puzzle and of its same-step successor
    // into L1 -- prefetch.global.L1 -- made the steady-state step slower, 8.30 -> 8.80 us on config 2: between graph
    // replays those few KB stay cached anyway.)
    if (tid < nE) {
        // a state that needs validating is left in shared memory: warp 3 validates it while warps 0-2 encode (inline,
        // the other 127 threads of the CTA waited at the barrier for it: 10.1 -> 8.8 us per step on config 2)
        int fl = 0;
        long long pk = SPARC_NO_ARG;
        if (packed_actions) { const long long i = env0 + tid; pk = (long long)((packed_actions[i >> 4] >> ((i & 15) * 2)) & 3u); }
        have_stats = phase_a<false, true>(p, env0 + tid, e, head + tid * SPARC_DYN_STRIDE, my_stats, &fl, snaps + tid * SPARC_VQ_WORDS, pk);
        tflags[tid] = (uint32_t)fl;
        head[p.tile * SPARC_DYN_STRIDE + tid] = e.puzzle;
        write_dyn(st, head, tid, e);
    }
    EP_MARK(0);
    if (p.mode == 0 && tid < ((nE + 31) & ~31)) reduce_stats(p, have_stats, my_stats);
    __syncthreads();
    EP_MARK(1);
    if (p.done_flag) {
        // sparc_step_host: every CTA counts itself once its per-env results (one packed byte each, device memory) are
        // written; the last one copies them to the host in one coalesced sweep and releases the host, which need not
        // wait for the observation rewrite that follows
        __shared__ int s_last;
        if (tid == 0) { __threadfence(); s_last = atomicAdd(p.done_count, 1u) == gridDim.x - 1; }
        __syncthreads();
        if (s_last) {
            __threadfence();
            const uint4* src = reinterpret_cast<const uint4*>(p.packed_out);
            for (int k = tid; k < (int)((n + 15) >> 4); k += CTA_THREADS) {       // 16 envs: 64 + 16 + 16 bytes over PCIe
                const uint4 q = __ldcg(src + k);                                  // (written by other SMs: not through L1)
                const uint32_t w[4] = {q.x, q.y, q.z, q.w};
                uint32_t te[4], tr[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    float4 r;
                    r.x = reward_value(w[u] & 7); r.y = reward_value((w[u] >> 8) & 7);
                    r.z = reward_value((w[u] >> 16) & 7); r.w = reward_value((w[u] >> 24) & 7);
                    reinterpret_cast<float4*>(p.host_reward)[4 * k + u] = r;
                    te[u] = (w[u] >> 3) & 0x01010101u; tr[u] = (w[u] >> 4) & 0x01010101u;
                }
                reinterpret_cast<uint4*>(p.host_term)[k] = make_uint4(te[0], te[1], te[2], te[3]);
                reinterpret_cast<uint4*>(p.host_trunc)[k] = make_uint4(tr[0], tr[1], tr[2], tr[3]);
            }
            __syncthreads();
            if (tid == 0) {
                *p.done_count = 0;
                __threadfence_system();
                *reinterpret_cast<volatile uint32_t*>(p.done_flag) = p.done_seq;
            }
        }
    }
    const int lane = tid & 31, warp = tid >> 5;
    uint32_t fl = 0;
    for (int j = lane; j < nE; j += 32) fl |= tflags[j];
    fl = __reduce_or_sync(0xffffffffu, fl);                       // the same in every warp of the CTA
    const bool any_snap = (fl & PA_SNAPSHOT) != 0;
    const bool any_obs = p.obs.base_dev || p.obs.color_dev || p.obs.info_dev;
    // persistent observation (update = SPARC_OBS_DELTA): phase A has touched the cells that changed; the tile is
    // rewritten only when one of its envs began a new episode (the others' in-place writes carry the same values)
    const bool encode = any_obs && !(p.mode == 0 && p.obs.update == SPARC_OBS_DELTA && !(fl & PA_NEW_EPISODE));
    if (any_snap && warp == CTA_THREADS / 32 - 1) {               // ---- the validator warp
        unsigned long long pass = 0;
        for (int j = lane; j < nE; j += 32)
            if (tflags[j] & PA_SNAPSHOT) {
                const uint32_t* q = snaps + j * SPARC_VQ_WORDS;
                pass += (unsigned long long)(validate_snapshot(p, q) & (int)q[3]);
            }
        pass = __reduce_add_sync(0xffffffffu, (unsigned)pass);
        if (lane == 0 && pass) atomicAdd(p.sd.stats + (size_t)(blockIdx.x % SPARC_STAT_ROWS) * SPARC_N_STATS + SPARC_STAT_RULES_PASS, pass);
        return;
    }
    if (!encode) return;
    const int nthr = any_snap ? CTA_THREADS - 32 : CTA_THREADS;   // the encoding threads
    encode_tile_stage1<FMT>(st, head, flats, p.tile, nE, p.obs, tid, nthr);
    if (any_snap) asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory"); else __syncthreads();
    EP_MARK(2);
    encode_tile_stage2<FMT>(st, head, flats, p.tile, nE, env0, p.obs, tid, nthr);
    EP_END();
}
template <int FMT>
__global__ void __launch_bounds__(CTA_THREADS, SPARC_MINBLOCKS) k_env(const KParams p) { k_env_body<FMT>(p, nullptr); }
template <int FMT>
__global__ void __launch_bounds__(CTA_THREADS, SPARC_MINBLOCKS) k_env_host(const KParams p, const PackedActions pa) { k_env_body<FMT>(p, pa.w); }

// ---- split pair for big batches: k_dyn (latency bound) || k_encode (bandwidth bound) -------------
// The observation is a pure function of the env state, so the step splits into a dynamics kernel
// (one env per thread, every lane busy) and an encoder kernel that streams state -> observation.
// sparc_step cuts the batch into chunks and runs k_dyn(chunk c+1) on a high-priority side stream
// while k_encode(chunk c) streams on the caller's stream: the dependent-load chains of the
// dynamics overlap the HBM-bound stores instead of sitting in front of them.
#ifndef DYN_MINBLOCKS
#define DYN_MINBLOCKS 1
#endif
// INLINE_VALIDATION = false: terminal states are only queued (k_val validates them), which drops the
// validator from this kernel's register budget (145 -> ~60) and doubles its resident warps.
// (Programmatic dependent launch of k_encode behind k_dyn -- griddepcontrol.launch_dependents at the top of k_dyn,
// griddepcontrol.wait in k_encode after a speculative prefetch of the puzzle's encoder inputs -- was measured inside
// the CUDA graph of bench.py: 0.196 -> 0.202 ms/step at 131,072 envs, 0.108 -> 0.117 at 65,536, 0.044 -> 0.051 at
// 16,384.  The parked encoder CTAs take registers and issue slots from the latency-bound dynamics; the graph's
// kernel-to-kernel edge is already ~1 us.  Removed.)
template <bool INLINE_VALIDATION, bool DELTA>
__global__ void __launch_bounds__(CTA_THREADS, DYN_MINBLOCKS) k_dyn(const KParams p, long long lo, long long hi) {
    const long long i = lo + (long long)blockIdx.x * CTA_THREADS + threadIdx.x;
    EnvRegs e;
    unsigned long long my_stats[SPARC_N_STATS];
    uint32_t scratch[SPARC_DYN_STRIDE];
    bool have_stats = false;
    if (i < hi) have_stats = phase_a<INLINE_VALIDATION, DELTA>(p, i, e, scratch, my_stats);
    if (p.mode == 0) reduce_stats(p, have_stats, my_stats);
}

// deferred terminal validation: the grid walks the queue k_dyn filled, one snapshot per WARP (validate_warp in
// sparc_val_warp.cuh: boards one word per lane, region rules one colour / one shape slot per lane, the exact-fit search
// split over the lanes).  Round 1 ran one snapshot per thread: 10.7 us of lone dependent chain each, 7.6 of 32 lanes
// active, the kernel as long as its slowest thread (24-39 us typical, 408 us at worst: profiles/r1_valprof.txt).
// Hard items: a snapshot whose exact-fit search outgrows the validating kernel's node budget (dense polyomino boards:
// 1e3..1e5 nodes; one node costs 0.1-0.8 us in one warp whatever its lanes do, so a big search needs many warps).
// Budgets and k_val_hard exist only for pools that can produce such searches (SparcEnv::hard_search: some puzzle holds
// eight or more shape instances, or SPARC_HARD_SEARCH=1): for other pools every search runs inline and k_val_hard is
// not launched at all -- its idle launch costs 2-3 us of every step, and an item costs more than an inline search of
// a few hundred nodes (measured, 131,072 envs of the config-4 pool, rewrite / compact step in ms: unlimited inline
// search 0.1825 / 0.0806, budget 768 + k_val_hard 0.1863 / 0.0859, budget 256 0.1891 / 0.0881).  Also measured and
// dropped: the items taken by lingering warps of the validating kernel itself (too few warps beside the encoder: p99
// 10.9 ms on dense boards) and a device-side tail launch of k_val_hard (-rdc slowed every kernel: step 0.186 -> 0.211 ms).
// hq_count words: 0 items, 1 shares handed out, 2 finished CTAs of k_val_hard.  hq item (uint4): queue slot, shares
// finished, found mask (64 bits).
__device__ __forceinline__ void push_hard_item(const KParams& p, uint32_t slot) {
    p.hq[atomicAdd(p.hq_count, 1u)] = make_uint4(slot, 0u, 0u, 0u);
}

// Terminal validation BESIDE the full-rewrite encoder: one snapshot per thread, a small grid.  k_encode fills the
// register file (8 CTAs x 60 registers x 128 threads per SM), so every validator CTA displaces encoder CTAs for as
// long as it runs; what counts here is snapshots per resident warp, not latency, and 32 diverging lanes that each
// carry a snapshot get through ~2.5x more of them per warp-second than a converged warp that takes them one by one
// (measured at 131,072 envs, ~2,230 snapshots per step, 32 CTAs: this kernel 20-50 us alone, the warp-per-snapshot
// one ~75 us alone and longer than the encoder beside it).
// Snapshots are dealt out warp by warp (slot s -> warp s mod nwarps): the fewer diverging lanes share a warp, the shorter.
__global__ void __launch_bounds__(CTA_THREADS) k_val(const KParams p) {
    const uint32_t count = *p.vq_count;
    unsigned long long pass = 0;
    const uint32_t nwarps = gridDim.x * (CTA_THREADS / 32);
    const uint32_t warp = blockIdx.x * (CTA_THREADS / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    for (uint32_t slot = lane * nwarps + warp; slot < count; slot += 32 * nwarps) {
        const int r = validate_queued(p, slot, p.fit_budget);
        if (r < 0) push_hard_item(p, slot);
        else pass += (unsigned long long)r;
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) pass += __shfl_down_sync(0xffffffffu, pass, off);
    if ((threadIdx.x & 31) == 0 && pass) atomicAdd(p.sd.stats + SPARC_STAT_RULES_PASS, pass);
    __syncthreads();                                             // every thread of the CTA has read `count`
    if (threadIdx.x == 0) {                                      // last CTA out empties the queue for the next step
        __threadfence();
        if (atomicAdd(p.vq_count + 1, 1u) == gridDim.x - 1) { p.vq_count[0] = 0; p.vq_count[1] = 0; }
    }
}

// Each hard item is validated again by VW_HARD_CODES shares, share c searching part c of every search tree
// (validate_warp<true> -> fit_step with code = c, the whole warp converged).  Shares are handed out through one
// counter, so the 32 shares of a snapshot run wherever a warp is free.  A share that finds a fit sets the region's bit
// in the item's found mask (and thereby stops its siblings); the last share of a snapshot to finish writes its flags.
// One warp per CTA: beside k_encode a 128-thread CTA of this kernel would need the registers of three retiring encoder CTAs.
#define HARD_CTAS 1184
__global__ void __launch_bounds__(32) k_val_hard(const KParams p) {
    const uint32_t n_items = p.hq_count[0];
    const int lane = threadIdx.x & 31;
    unsigned long long pass = 0;
    if (!n_items) return;
    const uint32_t total = n_items * VW_HARD_CODES;
    for (;;) {
        uint32_t w = 0;
        if (lane == 0) w = atomicAdd(p.hq_count + 1, 1u);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= total) break;
        uint4* item = p.hq + (w % n_items);                      // share-major: the snapshots start side by side
        const int code = (int)(w / n_items);
        const uint32_t slot = item->x;
        const uint32_t qw = lane < SPARC_VQ_WORDS ? p.vq[(size_t)slot * SPARC_VQ_WORDS + lane] : 0u;
        unsigned long long* found = reinterpret_cast<unsigned long long*>(&item->z);
        unsigned long long fit_mask = 0;
        int n_search = 0;
        int f = validate_warp<true>(p.st, qw, lane, nullptr, 0, code, found, &fit_mask, &n_search);
        if (lane == 0) {
            if (fit_mask) atomicOr(found, fit_mask);
            __threadfence();
            if (atomicAdd(&item->y, 1u) == VW_HARD_CODES - 1) {                                       // last share of this snapshot
                __threadfence();
                const unsigned long long all = atomicOr(found, 0ull);
                const unsigned long long need = n_search >= 64 ? ~0ull : (1ull << n_search) - 1ull;
                if ((all & need) != need) f &= ~SPARC_RULE_POLY_YLOP;                                // some region has no tiling
                if ((f & 0xFF) == 0xFF) f |= SPARC_RULE_ALL;
                if (p.out.rule_flags_dev) p.out.rule_flags_dev[qw] = (uint16_t)(f | SPARC_RULE_VALID);       // word 0 = env
                pass += (f & SPARC_RULE_ALL) != 0;
            }
        }
    }
    if (lane == 0) {
        if (pass) atomicAdd(p.sd.stats + SPARC_STAT_RULES_PASS, pass);
        __threadfence();                                         // last CTA out empties the hard queue
        if (atomicAdd(p.hq_count + 2, 1u) == gridDim.x - 1) { p.hq_count[0] = 0; p.hq_count[1] = 0; p.hq_count[2] = 0; }
    }
}

#ifndef ENC_THREADS
#define ENC_THREADS 128
#endif
#ifndef ENC_MINBLOCKS
#define ENC_MINBLOCKS 8
#endif
template <int FMT>
__global__ void __launch_bounds__(ENC_THREADS, ENC_MINBLOCKS) k_encode(const KParams p, long long lo, long long hi) {
    const StoreDev& st = p.st;
    const int tid = threadIdx.x;
    const long long env0 = lo + (long long)blockIdx.x * p.tile;
    const int nE = (int)min((long long)p.tile, hi - env0);
    uint32_t* head = smem;
    uint32_t* flats = smem + tile_head_words(p.tile);
    if (tid < nE) {
        EnvRegs e;
        load_env(e, p.sd, st, p.cfg.n_envs, env0 + tid);
        prefetch_static(st, e.puzzle);
        head[p.tile * SPARC_DYN_STRIDE + tid] = e.puzzle;
        write_dyn(st, head, tid, e);
    }
    __syncthreads();
    encode_tile_stage1<FMT>(st, head, flats, p.tile, nE, p.obs, tid, ENC_THREADS);
    __syncthreads();
    encode_tile_stage2<FMT>(st, head, flats, p.tile, nE, env0, p.obs, tid, ENC_THREADS);
}

// update = SPARC_OBS_DELTA on the split path: k_dyn touches the changed cells in place, queues the envs that began a new
// episode (1-2 % per step) and queues the terminal states; k_post does both follow-ups in ONE launch, one WARP per queue
// entry: first the warp-cooperative rule validation of its snapshots (validate_warp, sparc_val_warp.cuh: ~4 us each,
// the latency that bounds this step), then the full rewrite of its new-episode envs (rows of tile = 1 in the warp's own
// slice of shared memory, element-wise stores).  As two kernels on two streams they did not overlap: a grid that gives
// every snapshot its own warp fills the register file, so k_encode_list waited for k_val (step 0.046 ms = k_dyn 12 +
// k_val 12 + k_encode_list 11 + launches); fused, every warp has work of both kinds.
// Snapshots whose exact-fit search outgrows the node budget (pools with dense boards only) become hard items (k_val_hard).
template <int FMT>
__global__ void __launch_bounds__(ENC_THREADS) k_post(const KParams p) {
    __shared__ uint16_t fit_stack[ENC_THREADS / 32][128];
    const StoreDev& st = p.st;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t nwarps = gridDim.x * (ENC_THREADS / 32);
    const uint32_t warp = blockIdx.x * (ENC_THREADS / 32) + wib;
    unsigned long long pass = 0;
    if (p.vq_count) {                                            // ---- terminal validation
        const uint32_t count = *p.vq_count;
        for (uint32_t slot = warp; slot < count; slot += nwarps) {
            const uint32_t qw = lane < SPARC_VQ_WORDS ? p.vq[(size_t)slot * SPARC_VQ_WORDS + lane] : 0u;   // 48 contiguous bytes
            const int f = validate_warp<false>(p.st, qw, lane, fit_stack[wib], p.fit_budget);
            if (lane == 0) {
                if (f < 0) push_hard_item(p, slot);
                else {
                    if (p.out.rule_flags_dev) p.out.rule_flags_dev[qw] = (uint16_t)(f | SPARC_RULE_VALID); // word 0 = env
                    pass += (f & SPARC_RULE_ALL) != 0;
                }
            }
        }
    }
    const uint32_t rcount = p.rq_count ? *p.rq_count : 0u;      // ---- envs that began a new episode: whole observation
    const int per_warp = tile_head_words(1) + tile_flat_words(st, 1);
    uint32_t* head = smem + wib * per_warp;
    uint32_t* flats = head + tile_head_words(1);
    for (uint32_t slot = nwarps - 1 - warp; slot < rcount; slot += nwarps) {   // (the warps with one snapshot fewer go first)
        const long long env = (long long)p.rq[slot];
        if (lane == 0) {
            EnvRegs e;
            load_env(e, p.sd, st, p.cfg.n_envs, env);
            head[SPARC_DYN_STRIDE] = e.puzzle;
            write_dyn(st, head, 0, e);
        }
        __syncwarp();
        encode_env_stage1(st, head, flats, p.obs, lane, 32);
        __syncwarp();
        encode_env_elementwise<FMT>(st, head, flats, env, p.obs, lane, 32);
        __syncwarp();                                            // the rows are reused by the next pass
    }
    if (lane == 0 && pass) atomicAdd(p.sd.stats + SPARC_STAT_RULES_PASS, pass);
    __syncthreads();                                             // every thread of the CTA has read the counts
    if (threadIdx.x == 0) {                                      // last CTA out empties both queues for the next step
        __threadfence();
        if (atomicAdd(p.rq_count + 1, 1u) == gridDim.x - 1) {
            p.rq_count[0] = 0; p.rq_count[1] = 0;
            if (p.vq_count) { p.vq_count[0] = 0; p.vq_count[1] = 0; }
        }
    }
}

// stand-alone validator (SG.py:941-950) for sparc_validate: one env per thread
__global__ void __launch_bounds__(CTA_THREADS) k_validate(const StoreDev st, const StateDev sd, long long n,
                                                          const uint8_t* mask, uint16_t* flags, int8_t* region_map,
                                                          int8_t* region_fit) {
    const long long i = (long long)blockIdx.x * CTA_THREADS + threadIdx.x;
    if (i >= n) return;
    int8_t* rf = region_fit ? region_fit + (size_t)i * SPARC_MAX_REGIONS : nullptr;
    if (rf) for (int k = 0; k < SPARC_MAX_REGIONS; k++) rf[k] = SPARC_FIT_NONE;
    if (mask && !mask[i]) { if (flags) flags[i] = 0; return; }
    EnvRegs e;
    load_env(e, sd, st, n, i);
    uint32_t vd[9];
#pragma unroll
    for (int k = 0; k < 8; k++) vd[k] = e.v[k];
    vd[8] = 0;
    int8_t* rm = region_map ? region_map + (size_t)i * st.S : nullptr;
    if (rm) for (int k = 0; k < st.S; k++) rm[k] = -1;
    const int f = validate_env(st, e.puzzle, vd, e.x, e.y, rm, rf);
    if (flags) flags[i] = (uint16_t)(f | SPARC_RULE_VALID);
}

// sparc_step_info_host / sparc_reset_info_host for a handful of envs (the single-env Gymnasium drop-in): ONE launch does
// everything the call returns.  CTA = env: thread 0 steps (or resets) it -- the action / puzzle index came with the launch
// -- and validates the state it is left in, region map and per-region verdicts into shared memory; the CTA then writes
// the env's part of the host record (struct of arrays, see include/sparc_b200.h) with coalesced stores over PCIe, the
// last CTA raises the completion flag the host spins on, and the observation is rewritten after that.
// (As three launches -- k_env and k_validate writing into mapped host memory byte by byte, then a stream synchronise --
// one step of one env cost 37 us; tools/hostcall_probe.py.)
#define SPARC_INFO_MAX_ENVS 16
struct SmallInts { long long v[SPARC_INFO_MAX_ENVS]; };
template <int FMT>
__global__ void __launch_bounds__(CTA_THREADS) k_info(const KParams p, const SmallInts args, char* rec) {
    const StoreDev& st = p.st;
    const int tid = threadIdx.x;
    const long long i = blockIdx.x, n = p.cfg.n_envs;
    uint32_t* head = smem;                                       // tile = 1: dyn row + puzzle id | rows | this env's record
    uint32_t* flats = smem + tile_head_words(1);
    int8_t* rfit = reinterpret_cast<int8_t*>(flats + tile_flat_words(st, 1));
    int8_t* rmap = rfit + SPARC_MAX_REGIONS;
    __shared__ int s_small[8];
    __shared__ uint32_t s_snap[SPARC_VQ_WORDS];
    for (int k = tid; k < SPARC_MAX_REGIONS + st.S; k += CTA_THREADS) rfit[k] = k < SPARC_MAX_REGIONS ? (int8_t)SPARC_FIT_NONE : (int8_t)-1;
    __syncthreads();
    unsigned long long my_stats[SPARC_N_STATS];
    bool have_stats = false;
    if (tid == 0) {
        EnvRegs e;
        StepResult r;
        int fl = 0;
        uint32_t* snap = s_snap;
        have_stats = phase_a<false, true>(p, i, e, head, my_stats, &fl, snap, args.v[i], &r);
        if ((fl & PA_SNAPSHOT) && (validate_snapshot(p, snap) & (int)snap[3]))      // validate = terminal / every_step: the statistic
            atomicAdd(p.sd.stats + (size_t)(blockIdx.x % SPARC_STAT_ROWS) * SPARC_N_STATS + SPARC_STAT_RULES_PASS, 1ull);
        uint32_t vd[9];
#pragma unroll
        for (int k = 0; k < 8; k++) vd[k] = e.v[k];
        vd[8] = 0;
        const int f = validate_env(st, e.puzzle, vd, e.x, e.y, rmap, rfit) | SPARC_RULE_VALID;      // SG.py:1011: the state as it is now
        s_small[0] = (int)e.step; s_small[1] = f; s_small[2] = e.x | (e.y << 8);
        s_small[3] = (r.reward_code & 0xff) | (r.terminated << 8) | (r.truncated << 16) | (r.legal_mask << 24);
        head[SPARC_DYN_STRIDE] = e.puzzle;
        write_dyn(st, head, 0, e);
    }
    if (p.mode == 0 && tid < 32) reduce_stats(p, have_stats, my_stats);
    __syncthreads();
    if (tid == 0) {                                              // int32 step[n] | uint16 flags[n] | uint8 agent_xy[n][2] | int8 reward_code[n] |
        reinterpret_cast<int32_t*>(rec)[i] = s_small[0];        // uint8 terminated[n] | uint8 truncated[n] | uint8 legal_mask[n]
        reinterpret_cast<uint16_t*>(rec + 4 * n)[i] = (uint16_t)s_small[1];
        reinterpret_cast<uint16_t*>(rec + 6 * n)[i] = (uint16_t)s_small[2];
        rec[8 * n + i] = (char)(s_small[3] & 0xff); rec[9 * n + i] = (char)((s_small[3] >> 8) & 0xff);
        rec[10 * n + i] = (char)((s_small[3] >> 16) & 0xff); rec[11 * n + i] = (char)((s_small[3] >> 24) & 0xff);
    }
    int8_t* fit_out = reinterpret_cast<int8_t*>(rec + 12 * n) + (size_t)i * SPARC_MAX_REGIONS;    // | int8 region_fit[n][SPARC_MAX_REGIONS]
    int8_t* map_out = reinterpret_cast<int8_t*>(rec + (12 + SPARC_MAX_REGIONS) * n) + (size_t)i * st.S;   // | int8 region_map[n][S]
    for (int k = tid; k < SPARC_MAX_REGIONS; k += CTA_THREADS) fit_out[k] = rfit[k];
    for (int k = tid; k < st.S; k += CTA_THREADS) map_out[k] = rmap[k];
    __syncthreads();
    if (tid == 0) {
        __threadfence_system();
        if (atomicAdd(p.done_count, 1u) == gridDim.x - 1) {
            *p.done_count = 0;
            __threadfence_system();
            *reinterpret_cast<volatile uint32_t*>(p.done_flag) = p.done_seq;
        }
    }
    if (!p.obs.base_dev && !p.obs.color_dev && !p.obs.info_dev) return;
    encode_env_stage1(st, head, flats, p.obs, tid, CTA_THREADS);
    __syncthreads();
    encode_env_elementwise<FMT>(st, head, flats, i, p.obs, tid, CTA_THREADS);
}

// a fresh env is a clean reset on puzzle i mod pool: stepping, encoding or validating before the first sparc_reset
// works on a state a puzzle can produce (start node visited, agent on the start)
__global__ void k_init_state(StoreDev st, StateDev sd, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    EnvRegs e;
    env_reset(e, st, (uint32_t)(i % st.n_puzzles));
    store_env(e, sd, st, n, i);
}

__global__ void k_read_stats(unsigned long long* stats, long long* out, int clear) {     // <<<1, 32 * SPARC_N_STATS>>>
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;                            // warp k sums counter k over the rows
    unsigned long long v = 0;
    for (int r = lane; r < SPARC_STAT_ROWS; r += 32) {
        v += stats[(size_t)r * SPARC_N_STATS + k];
        if (clear) stats[(size_t)r * SPARC_N_STATS + k] = 0;
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if (lane == 0) out[k] = (long long)v;
}

// Every entry point works on the store's device and leaves the calling thread's current device as it found it (a
// caller with several GPUs in one process -- or torch, which reads the current device back -- must not see it move).
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev); else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
template <typename T>
static int upload(const T* host, size_t count, const T** dev, void** slot) {
    void* d = nullptr;
    const size_t bytes = (count ? count : 1) * sizeof(T);
    CU(cudaMalloc(&d, bytes));
    *slot = d;
    if (count) CU(cudaMemcpy(d, host, count * sizeof(T), cudaMemcpyHostToDevice));
    *dev = reinterpret_cast<const T*>(d);
    return SPARC_OK;
}

extern "C" {

int sparc_abi_version(void) { return SPARC_ABI_VERSION; }
const char* sparc_last_error(void) { return g_err; }

int sparc_store_create(const SparcStoreDesc* desc, int device, SparcStore** out) {
    if (!desc || !out) return fail(SPARC_E_ARG, "null argument%s");
    if (desc->n_puzzles <= 0 || desc->slab_x <= 0 || desc->slab_x > SPARC_MAX_DIM || desc->slab_y <= 0 ||
        desc->slab_y > SPARC_MAX_DIM || desc->n_planes < 4 || desc->n_planes > SPARC_MAX_PLANES)
        return fail(SPARC_E_ARG, "bad store geometry%s");
    if (desc->plane_ids[0] != SPARC_PL_VISITED || desc->plane_ids[2] != SPARC_PL_AGENT)
        return fail(SPARC_E_ARG, "planes 0 and 2 must be visited and agent_location%s");
    DeviceGuard dg_(device); CU(dg_.err);
    SparcStore* s = new (std::nothrow) SparcStore();
    if (!s) return fail(SPARC_E_NOMEM, "out of host memory%s");
    memset(s, 0, sizeof(*s));
    s->device = device;
    StoreDev& d = s->d;
    d.n_puzzles = desc->n_puzzles; d.SX = desc->slab_x; d.SY = desc->slab_y; d.S = d.SX * d.SY;
    d.P = desc->n_planes; d.L = d.P * d.S; d.W = (d.L + 31) / 32;
    if (desc->words_per_puzzle != d.W) { delete s; return fail(SPARC_E_ARG, "words_per_puzzle mismatch%s"); }
    if ((uint64_t)desc->n_puzzles * (uint64_t)(d.W + 2) >= (1ull << 32)) {        // the encoder indexes the pool with 32-bit word offsets
        delete s; return fail(SPARC_E_ARG, "pool too large for 32-bit word offsets%s");
    }
    d.nv = (d.S + 31) / 32;
    d.inv_L = (uint32_t)((1ull << 32) / (uint64_t)d.L) + 1u;
    d.inv_S = (uint32_t)((1ull << 32) / (uint64_t)d.S) + 1u;
    d.inv_4S = (uint32_t)((1ull << 32) / (uint64_t)(4 * d.S)) + 1u;
    d.Wc = (4 * d.S + 31) / 32;
    { const uint64_t k0b = (uint64_t)(d.L >> 5), k0c = (uint64_t)((4 * d.S) >> 5), k0z = (uint64_t)(d.S >> 5);
      d.inv_k0b = k0b ? (uint32_t)((1ull << 32) / k0b) + 1u : 0u; d.inv_k0c = k0c ? (uint32_t)((1ull << 32) / k0c) + 1u : 0u;
      d.inv_k0z = k0z ? (uint32_t)((1ull << 32) / k0z) + 1u : 0u; }
    memcpy(s->plane_ids, desc->plane_ids, sizeof(s->plane_ids));
    const size_t n = desc->n_puzzles;
    int rc;
    if ((rc = upload(desc->headers, n, &d.hdr, &s->blocks[0])) || (rc = upload(desc->node_boards, n * SPARC_N_NODE_BOARDS * 8, &d.nb, &s->blocks[1])) ||
        (rc = upload(desc->cell_boards, n * SPARC_N_CELL_BOARDS, &d.cb, &s->blocks[2])) || (rc = upload(desc->static_bits, n * d.W, &d.sbits, &s->blocks[3])) ||
        (rc = upload(desc->color4, n * d.Wc, &d.color4, &s->blocks[4])) || (rc = upload(desc->info_nz, n * d.nv, &d.info_nz, &s->blocks[8])) || (rc = upload(desc->info, n * d.S, &d.info, &s->blocks[5])) ||
        (rc = upload(desc->shapes, (size_t)desc->n_shapes_total, &d.shapes, &s->blocks[6])) ||
        (rc = upload(desc->trie, (size_t)desc->n_trie_nodes * SPARC_TRIE_NODE_WORDS, &d.trie, &s->blocks[7]))) {
        sparc_store_destroy(s);
        return rc;
    }
    for (size_t i = 0; i < n; i++) {
        int inst = 0;
        for (int k = 0; k < desc->headers[i].n_shapes; k++) {
            const SparcShape& sh = desc->shapes[desc->headers[i].shape_off + k];
            inst += __builtin_popcountll(sh.poly_cells) + __builtin_popcountll(sh.ylop_cells);
        }
        if (inst > s->max_instances) s->max_instances = inst;
    }
    {   // validator records of k_val (layout: sparc_val_warp.cuh), assembled from the arrays above
        uint32_t* vr = new (std::nothrow) uint32_t[n * SPARC_VREC_WORDS];
        if (!vr) { sparc_store_destroy(s); return fail(SPARC_E_NOMEM, "out of host memory%s"); }
        static const int nb_of[9] = {SPARC_NB_GAPS16, SPARC_NB_CELL16, SPARC_NB_VALID16, SPARC_NB_DOT16, SPARC_NB_TRI1,
                                     SPARC_NB_TRI2, SPARC_NB_TRI3, SPARC_NB_TRI4, SPARC_NB_TRIBAD};
        for (size_t i = 0; i < n; i++) {
            const SparcPuzzleHeader& h = desc->headers[i];
            const uint32_t* nb = desc->node_boards + i * SPARC_N_NODE_BOARDS * 8;
            const uint64_t* cb = desc->cell_boards + i * SPARC_N_CELL_BOARDS;
            uint64_t anycol = 0;
            for (int c = 0; c < 8; c++) anycol |= cb[SPARC_CB_COLOR1 + c];
            for (int k = 0; k < 8; k++) {
                uint32_t* r = vr + (i * 8 + k) * SPARC_VREC_ROW_WORDS;
                for (int b = 0; b < 9; b++) r[b] = nb[nb_of[b] * 8 + k];
                r[9] = (uint32_t)h.X | ((uint32_t)h.Y << 8) | ((uint32_t)h.start_x << 16) | ((uint32_t)h.start_y << 24);
                r[10] = (uint32_t)h.target_x | ((uint32_t)h.target_y << 8) | ((uint32_t)h.n_shapes << 16) | ((uint32_t)h.flags << 24);
                r[11] = h.shape_off;
                const uint64_t cells[6] = {cb[SPARC_CB_STAR], cb[SPARC_CB_SQUARE], cb[SPARC_CB_SYM1], cb[SPARC_CB_SYM2],
                                           cb[SPARC_CB_COLOR1 + k], anycol};
                for (int c = 0; c < 6; c++) { r[12 + 2 * c] = (uint32_t)cells[c]; r[13 + 2 * c] = (uint32_t)(cells[c] >> 32); }
            }
        }
        rc = upload(vr, n * SPARC_VREC_WORDS, &d.vrec, &s->blocks[9]);
        delete[] vr;
        if (rc) { sparc_store_destroy(s); return rc; }
    }
    *out = s;
    return SPARC_OK;
}

int sparc_store_destroy(SparcStore* s) {
    if (!s) return SPARC_OK;
    DeviceGuard dg_(s->device);
    for (void* b : s->blocks) if (b) cudaFree(b);
    delete s;
    return SPARC_OK;
}

// uint8 streams stay 16-byte aligned only when every tile starts at a multiple of 16 envs
static int compact_tile(int t) { t = (t + 15) / 16 * 16; return t > CTA_THREADS ? CTA_THREADS : t; }

extern "C" int sparc_env_destroy(SparcEnv* e);

static int pick_tile(int64_t n) {
    const char* env = getenv("SPARC_TILE");
    if (env) { int t = atoi(env); if (t >= 4 && t <= CTA_THREADS && t % 4 == 0) return t; }
    // largest tile that still gives every SM several CTAs (148 SMs x 8 CTAs)
    for (int t = 32; t > 4; t >>= 1) if (n / t >= 148 * 8) return t;
    return 4;
}

int sparc_env_create(SparcStore* store, int64_t n_envs, int traceback, int max_steps, int autoreset_mode,
                     int puzzle_stride, int validate_mode, SparcEnv** out) {
    if (!store || !out || n_envs <= 0) return fail(SPARC_E_ARG, "bad argument%s");
    if (autoreset_mode < 0 || autoreset_mode > 3 || validate_mode < 0 || validate_mode > 2 || puzzle_stride < 0)
        return fail(SPARC_E_ARG, "bad mode%s");
    DeviceGuard dg_(store->device); CU(dg_.err);
    SparcEnv* e = new (std::nothrow) SparcEnv();
    if (!e) return fail(SPARC_E_NOMEM, "out of host memory%s");
    memset(e, 0, sizeof(*e));
    e->store = store;
    e->cfg.n_envs = n_envs; e->cfg.traceback = traceback != 0; e->cfg.max_steps = max_steps;
    e->cfg.autoreset = autoreset_mode; e->cfg.stride = puzzle_stride; e->cfg.validate_mode = validate_mode;
    e->tile = pick_tile(n_envs);
    {   // big batches: the k_dyn || k_encode pipeline
        const char* pe = getenv("SPARC_SPLIT");                // chunks; 0 = fused k_env; unset = automatic
        int nc = pe ? atoi(pe) : (n_envs >= 8192 ? 1 : 0);    // measured crossover with the fused k_env: ~6 k envs;
                                                               // 1 chunk (k_dyn, then k_val || k_encode) beats 2..8 chunks
        if (nc > 16) nc = 16;
        const char* te = getenv("SPARC_ENC_TILE");
        e->enc_tile = te ? atoi(te) : 8;
        if (e->enc_tile < 4 || e->enc_tile > ENC_THREADS || e->enc_tile % 4) e->enc_tile = 8;
        {   // every tile a launch can pick (fused / split, reference / compact) must fit the kernels' dynamic shared memory
            int tmax = e->tile > e->enc_tile ? e->tile : e->enc_tile;
            if (tmax < 32) tmax = 32;                               // compact default
            tmax = compact_tile(tmax);
            const size_t shmem = (size_t)(tile_head_words(tmax) + tile_flat_words(store->d, tmax) + tile_snap_words(tmax)) * sizeof(uint32_t);
            if (shmem > 48 * 1024) {
                cudaError_t ae = cudaFuncSetAttribute(k_env<SPARC_OBS_REFERENCE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem);
                if (ae == cudaSuccess) ae = cudaFuncSetAttribute(k_env<SPARC_OBS_COMPACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem);
                if (ae == cudaSuccess) ae = cudaFuncSetAttribute(k_env_host<SPARC_OBS_REFERENCE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem);
                if (ae == cudaSuccess) ae = cudaFuncSetAttribute(k_env_host<SPARC_OBS_COMPACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem);
                if (ae == cudaSuccess) ae = cudaFuncSetAttribute(k_encode<SPARC_OBS_REFERENCE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem);
                if (ae == cudaSuccess) ae = cudaFuncSetAttribute(k_encode<SPARC_OBS_COMPACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem);
                if (ae != cudaSuccess) { sparc_env_destroy(e); return fail(SPARC_E_CUDA, "dynamic shared memory: %s", cudaGetErrorString(ae)); }
            }
        }
        if (nc > 0) {
            int lo_p = 0, hi_p = 0;
            cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p);
            cudaError_t se = cudaStreamCreateWithPriority(&e->side, cudaStreamNonBlocking, hi_p);
            if (se == cudaSuccess) se = cudaEventCreateWithFlags(&e->ev_start, cudaEventDisableTiming);
            for (int c = 0; c < nc && se == cudaSuccess; c++) se = cudaEventCreateWithFlags(&e->ev_dyn[c], cudaEventDisableTiming);
            if (se == cudaSuccess) se = cudaEventCreateWithFlags(&e->ev_val, cudaEventDisableTiming);
            if (se == cudaSuccess) se = cudaStreamCreateWithPriority(&e->copy_stream, cudaStreamNonBlocking, hi_p);
            if (se == cudaSuccess) se = cudaEventCreateWithFlags(&e->ev_dyn_done, cudaEventDisableTiming);
            if (se == cudaSuccess) se = cudaEventCreateWithFlags(&e->ev_copy, cudaEventDisableTiming);
            if (se == cudaSuccess) se = cudaEventCreateWithFlags(&e->ev_act, cudaEventDisableTiming);
            const char* de = getenv("SPARC_DEFER_VALIDATION");
            if (se == cudaSuccess && !(de && atoi(de) == 0)) {
                se = cudaMalloc((void**)&e->vq_count, 4 * sizeof(uint32_t));   // see k_val
                if (se == cudaSuccess) se = cudaMalloc((void**)&e->vq, (size_t)n_envs * SPARC_VQ_WORDS * sizeof(uint32_t));
                if (se == cudaSuccess) se = cudaMemset(e->vq_count, 0, 4 * sizeof(uint32_t));
                if (se == cudaSuccess) se = cudaMalloc((void**)&e->hq_count, 4 * sizeof(uint32_t));
                if (se == cudaSuccess) se = cudaMalloc((void**)&e->hq, (size_t)n_envs * sizeof(uint4));
                if (se == cudaSuccess) se = cudaMemset(e->hq_count, 0, 4 * sizeof(uint32_t));
                const char* he = getenv("SPARC_HARD_SEARCH");                  // 0 / 1; unset = by the pool's densest puzzle
                e->hard_search = he ? atoi(he) != 0 : store->max_instances >= 8;
                e->val_grid = 32;              // x128 threads; a small grid disturbs the co-running encoder least (measured)
                const char* ve = getenv("SPARC_VAL_GRID");
                if (ve && atoi(ve) > 0) e->val_grid = atoi(ve);
                e->val_grid_delta = 592;       // k_post: at most one resident wave of warps, one queue entry each
                const char* vd = getenv("SPARC_VAL_GRID_DELTA");
                if (vd && atoi(vd) > 0) e->val_grid_delta = atoi(vd);

            }
            if (se != cudaSuccess) { sparc_env_destroy(e); return fail(SPARC_E_CUDA, "side stream / queue: %s", cudaGetErrorString(se)); }
            e->split_chunks = nc;
        }
    }
    const size_t n = (size_t)n_envs;
    cudaError_t err = cudaSuccess;
    auto alloc = [&](void** p, size_t bytes) { if (err == cudaSuccess) { err = cudaMalloc(p, bytes); if (err == cudaSuccess) err = cudaMemset(*p, 0, bytes); } };
    alloc((void**)&e->s.core, n * 16);
    alloc((void**)&e->s.vis, n * 8 * 4);
    alloc((void**)&e->s.ep_ret, n * 4);
    if (traceback) alloc((void**)&e->s.moves, n * SPARC_MOVE_WORDS * 4);
    alloc((void**)&e->s.stats, (size_t)SPARC_STAT_ROWS * SPARC_N_STATS * 8);
    alloc((void**)&e->rq_count, 2 * sizeof(uint32_t));
    alloc((void**)&e->rq, n * sizeof(uint32_t));
    alloc((void**)&e->act_stage, n);
    alloc((void**)&e->rew_stage, n * 4);
    alloc((void**)&e->term_stage, n);
    alloc((void**)&e->trunc_stage, n);
    if (n <= SPARC_PACKED_MAX_ENVS) {                            // sparc_step_host on launch-bound batches (k_env_host)
        const size_t nb = (n + 15) / 16 * 16;
        alloc((void**)&e->done_count, sizeof(uint32_t));
        alloc((void**)&e->packed_dev, nb);
        if (err == cudaSuccess) err = cudaHostAlloc((void**)&e->done_host, 64, cudaHostAllocMapped);
        if (err == cudaSuccess) { *e->done_host = 0; err = cudaHostGetDevicePointer((void**)&e->done_dev, (void*)e->done_host, 0); }
        e->res_nb = nb;
        if (err == cudaSuccess) err = cudaHostAlloc((void**)&e->res_host, nb * 6, cudaHostAllocMapped);
        if (err == cudaSuccess) err = cudaHostGetDevicePointer((void**)&e->res_host_dev, (void*)e->res_host, 0);
    }
    if (err != cudaSuccess) { sparc_env_destroy(e); return fail(SPARC_E_CUDA, "cudaMalloc: %s", cudaGetErrorString(err)); }
    k_init_state<<<(unsigned)((n + 255) / 256), 256>>>(store->d, e->s, n_envs);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    *out = e;
    return SPARC_OK;
}

int sparc_env_destroy(SparcEnv* e) {
    if (!e) return SPARC_OK;
    DeviceGuard dg_(e->store->device);
    void* ptrs[] = {e->s.core, e->s.vis, e->s.ep_ret, e->s.moves, e->s.stats, e->vq_count, e->vq, e->hq_count, e->hq, e->rq_count, e->rq, e->act_stage, e->rew_stage, e->term_stage, e->trunc_stage, e->info_stage, e->int_stage, e->done_count, e->packed_dev};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (e->done_host) cudaFreeHost((void*)e->done_host);
    if (e->res_host) cudaFreeHost(e->res_host);
    if (e->info_host) cudaFreeHost(e->info_host);
    if (e->side) cudaStreamDestroy(e->side);
    if (e->ev_start) cudaEventDestroy(e->ev_start);
    for (cudaEvent_t ev : e->ev_dyn) if (ev) cudaEventDestroy(ev);
    if (e->ev_val) cudaEventDestroy(e->ev_val);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    if (e->ev_dyn_done) cudaEventDestroy(e->ev_dyn_done);
    if (e->ev_copy) cudaEventDestroy(e->ev_copy);
    if (e->ev_act) cudaEventDestroy(e->ev_act);
    delete e;
    return SPARC_OK;
}

static int launch_env(SparcEnv* e, KParams& p, void* stream) {
    DeviceGuard dg_(e->store->device); CU(dg_.err);   // raw C-ABI callers on multi-GPU hosts need not select the device themselves
    p.st = e->store->d; p.sd = e->s; p.cfg = e->cfg; p.tile = e->tile;
    const int64_t n = e->cfg.n_envs;
    if (p.obs.format != SPARC_OBS_REFERENCE && p.obs.format != SPARC_OBS_COMPACT) return fail(SPARC_E_ARG, "bad observation format%s");
    const bool compact = p.obs.format == SPARC_OBS_COMPACT;
    if (compact) p.tile = compact_tile(p.tile);
    const unsigned ntiles = (unsigned)((n + p.tile - 1) / p.tile);
    if (p.obs.update != SPARC_OBS_REWRITE && p.obs.update != SPARC_OBS_DELTA) return fail(SPARC_E_ARG, "bad observation update mode%s");
    if (p.mode == 1) p.obs.update = SPARC_OBS_REWRITE;                   // a reset always writes in full
    if (e->split_chunks > 0) {
        // chunk c: [k_dyn on the side stream] -> event -> [k_encode on the caller's stream]
        cudaStream_t s0 = (cudaStream_t)stream;
        const bool any_obs = p.obs.base_dev || p.obs.color_dev || p.obs.info_dev;
        const int nc = e->split_chunks;
        const int64_t per = ((n + nc - 1) / nc + 127) / 128 * 128;       // chunk boundaries stay 16-B aligned in every stream
        // compact: 32 envs = 81 KB per CTA unless SPARC_ENC_TILE says otherwise (measured with plain-row staging,
        // k_encode / step in us: 16 -> 59.7 / 86.7, 32 -> 59.8 / 82.3, 48 -> 62.6 / 83.8, 64 -> 67.9 / 87.2, 96 -> 80.7 / 95.5)
        p.tile = compact ? (getenv("SPARC_ENC_TILE") ? compact_tile(e->enc_tile) : 32) : e->enc_tile;
        const size_t shmem = (size_t)(tile_head_words(p.tile) + tile_flat_words(p.st, p.tile)) * sizeof(uint32_t);
        cudaStream_t sd = nc > 1 ? e->side : s0;                         // one chunk: plain back-to-back launches
        const bool defer = p.mode == 0 && e->vq && e->cfg.validate_mode == SPARC_VALIDATE_TERMINAL;
        const bool delta = any_obs && p.obs.update == SPARC_OBS_DELTA;
        if (delta) { p.rq = e->rq; p.rq_count = e->rq_count; }          // (k_post leaves the queue empty)
        if (defer) {
            p.vq = e->vq; p.vq_count = e->vq_count;                     // (k_val leaves the queue empty)
            p.hq = e->hq; p.hq_count = e->hq_count;
            // node budget of an inline exact-fit search (0 = none): one thread (k_val) / one converged warp (k_post)
            p.fit_budget = e->hard_search ? (delta ? 128 : 192) : 0;
        }
        if (nc > 1) {
            CU(cudaEventRecord(e->ev_start, s0));
            CU(cudaStreamWaitEvent(e->side, e->ev_start, 0));
        }
        for (int c = 0; c < nc; c++) {
            const long long lo = (long long)c * per, hi = lo + per < n ? lo + per : n;
            if (lo >= hi) break;
            const unsigned dgrid = (unsigned)((hi - lo + CTA_THREADS - 1) / CTA_THREADS);
            if (defer) { if (delta) k_dyn<false, true><<<dgrid, CTA_THREADS, 0, sd>>>(p, lo, hi); else k_dyn<false, false><<<dgrid, CTA_THREADS, 0, sd>>>(p, lo, hi); }
            else { if (delta) k_dyn<true, true><<<dgrid, CTA_THREADS, 0, sd>>>(p, lo, hi); else k_dyn<true, false><<<dgrid, CTA_THREADS, 0, sd>>>(p, lo, hi); }
            if (nc > 1) {
                CU(cudaEventRecord(e->ev_dyn[c], e->side));
                CU(cudaStreamWaitEvent(s0, e->ev_dyn[c], 0));
            }
            const bool last = hi >= n;
            const bool fork_val = defer && last && !delta;                   // (delta: k_post validates)
            const bool fork_copy = e->hostcopy.active && nc == 1;
            if (nc == 1 && (fork_val || fork_copy)) CU(cudaEventRecord(e->ev_start, s0));   // = k_dyn done
            if (delta) {
                if (last) {      // terminal validation + the envs that began a new episode in full: one warp per queue entry
                    const size_t shmem1 = (size_t)(ENC_THREADS / 32) * (tile_head_words(1) + tile_flat_words(p.st, 1)) * sizeof(uint32_t);
                    // ~1.6 % of the envs end an episode per step under random play: a warp per entry when the grid's warps
                    // cover n / 55 entries; at most one resident wave (4 CTAs per SM at k_post's registers)
                    int64_t want = (n + 219) / 220;
                    if (want > e->val_grid_delta) want = e->val_grid_delta;
                    const unsigned pgrid = (unsigned)(want < 1 ? 1 : want);
                    if (compact) k_post<SPARC_OBS_COMPACT><<<pgrid, ENC_THREADS, shmem1, s0>>>(p);
                    else k_post<SPARC_OBS_REFERENCE><<<pgrid, ENC_THREADS, shmem1, s0>>>(p);
                    if (defer && e->hard_search) { k_val_hard<<<HARD_CTAS, 32, 0, s0>>>(p); e->launches++; }
                }
            } else if (any_obs) {
                const unsigned egrid = (unsigned)((hi - lo + p.tile - 1) / p.tile);
                if (compact) k_encode<SPARC_OBS_COMPACT><<<egrid, ENC_THREADS, shmem, s0>>>(p, lo, hi);
                else k_encode<SPARC_OBS_REFERENCE><<<egrid, ENC_THREADS, shmem, s0>>>(p, lo, hi);
            }
            if (fork_copy) {                                                 // results -> host beside k_encode
                const size_t nn = (size_t)n;
                CU(cudaStreamWaitEvent(e->copy_stream, e->ev_start, 0));
                if (e->hostcopy.reward) CU(cudaMemcpyAsync(e->hostcopy.reward, e->rew_stage, nn * 4, cudaMemcpyDeviceToHost, e->copy_stream));
                if (e->hostcopy.term) CU(cudaMemcpyAsync(e->hostcopy.term, e->term_stage, nn, cudaMemcpyDeviceToHost, e->copy_stream));
                if (e->hostcopy.trunc) CU(cudaMemcpyAsync(e->hostcopy.trunc, e->trunc_stage, nn, cudaMemcpyDeviceToHost, e->copy_stream));
                CU(cudaEventRecord(e->ev_copy, e->copy_stream));
            }
            // (Launching k_val BEFORE k_encode -- so that its 124-register CTAs need not wait for encoder CTAs to retire
            // from a full register file -- changed nothing: compact 0.0833 vs 0.0833 ms, reference 0.1861 vs 0.1850.)
            if (fork_val) {                                                  // k_val beside (the last) k_encode
                if (nc == 1) CU(cudaStreamWaitEvent(e->side, e->ev_start, 0));
                k_val<<<(unsigned)e->val_grid, CTA_THREADS, 0, e->side>>>(p);
                if (e->hard_search) { k_val_hard<<<HARD_CTAS, 32, 0, e->side>>>(p); e->launches++; }
                CU(cudaEventRecord(e->ev_val, e->side));
                e->launches++;
                CU(cudaStreamWaitEvent(s0, e->ev_val, 0));                   // join
            }
            if (fork_copy) CU(cudaStreamWaitEvent(s0, e->ev_copy, 0));
            e->launches += (delta ? last : any_obs) ? 2 : 1;
        }
        CU(cudaGetLastError());
        return SPARC_OK;
    }
    {
        const size_t shmem = (size_t)(tile_head_words(p.tile) + tile_flat_words(p.st, p.tile) + tile_snap_words(p.tile)) * sizeof(uint32_t);
        if (e->notify) {
            p.done_flag = e->done_dev; p.done_count = e->done_count; p.done_seq = e->done_seq;
            p.packed_out = e->packed_dev;
            p.host_reward = reinterpret_cast<float*>(e->res_host_dev);
            p.host_term = reinterpret_cast<uint8_t*>(e->res_host_dev + 4 * e->res_nb);
            p.host_trunc = reinterpret_cast<uint8_t*>(e->res_host_dev + 5 * e->res_nb);
            if (compact) k_env_host<SPARC_OBS_COMPACT><<<ntiles, CTA_THREADS, shmem, (cudaStream_t)stream>>>(p, e->pa);
            else k_env_host<SPARC_OBS_REFERENCE><<<ntiles, CTA_THREADS, shmem, (cudaStream_t)stream>>>(p, e->pa);
        } else if (compact) k_env<SPARC_OBS_COMPACT><<<ntiles, CTA_THREADS, shmem, (cudaStream_t)stream>>>(p);
        else k_env<SPARC_OBS_REFERENCE><<<ntiles, CTA_THREADS, shmem, (cudaStream_t)stream>>>(p);
    }
    CU(cudaGetLastError());
    e->launches++;
    return SPARC_OK;
}

int sparc_reset(SparcEnv* e, const int32_t* puzzle_idx_dev, const uint8_t* mask_dev, const SparcStepOut* info_out,
                const SparcObsOut* obs, void* stream) {
    NvtxRange nvtx_("sparc_reset");
    if (!e) return fail(SPARC_E_ARG, "null env%s");
    KParams p; memset(&p, 0, sizeof(p));
    p.mode = 1; p.reset_idx = puzzle_idx_dev; p.reset_mask = mask_dev;
    if (info_out) p.out = *info_out;
    if (obs) p.obs = *obs;
    return launch_env(e, p, stream);
}

int sparc_step(SparcEnv* e, const void* actions_dev, int action_dtype, const SparcStepOut* out, const SparcObsOut* obs, void* stream) {
    NvtxRange nvtx_("sparc_step");
    if (!e || !actions_dev) return fail(SPARC_E_ARG, "null argument%s");
    if (action_dtype < 0 || action_dtype > 2) return fail(SPARC_E_ARG, "bad action dtype%s");
    KParams p; memset(&p, 0, sizeof(p));
    p.mode = 0; p.actions = actions_dev; p.act_dtype = action_dtype;
    if (out) p.out = *out;
    if (obs) p.obs = *obs;
    return launch_env(e, p, stream);
}

int sparc_encode(SparcEnv* e, const SparcObsOut* obs, void* stream) {
    NvtxRange nvtx_("sparc_encode");
    if (!e || !obs) return fail(SPARC_E_ARG, "null argument%s");
    DeviceGuard dg_(e->store->device); CU(dg_.err);
    KParams p; memset(&p, 0, sizeof(p));
    p.st = e->store->d; p.sd = e->s; p.cfg = e->cfg; p.obs = *obs;
    const int64_t n = e->cfg.n_envs;
    if (obs->format != SPARC_OBS_REFERENCE && obs->format != SPARC_OBS_COMPACT) return fail(SPARC_E_ARG, "bad observation format%s");
    p.tile = e->split_chunks > 0 ? e->enc_tile : (e->tile <= ENC_THREADS ? e->tile : 32);
    if (obs->format == SPARC_OBS_COMPACT) p.tile = p.tile < 32 ? 32 : compact_tile(p.tile);
    const size_t shmem = (size_t)(tile_head_words(p.tile) + tile_flat_words(p.st, p.tile)) * sizeof(uint32_t);
    const unsigned egrid = (unsigned)((n + p.tile - 1) / p.tile);
    if (obs->format == SPARC_OBS_COMPACT) k_encode<SPARC_OBS_COMPACT><<<egrid, ENC_THREADS, shmem, (cudaStream_t)stream>>>(p, 0, n);
    else k_encode<SPARC_OBS_REFERENCE><<<egrid, ENC_THREADS, shmem, (cudaStream_t)stream>>>(p, 0, n);
    CU(cudaGetLastError());
    e->launches++;
    return SPARC_OK;
}

int sparc_validate(SparcEnv* e, const uint8_t* mask_dev, uint16_t* flags_dev, int8_t* region_map_dev,
                   int8_t* region_fit_dev, void* stream) {
    NvtxRange nvtx_("sparc_validate");
    if (!e) return fail(SPARC_E_ARG, "null env%s");
    DeviceGuard dg_(e->store->device); CU(dg_.err);
    const int64_t n = e->cfg.n_envs;
    k_validate<<<(unsigned)((n + CTA_THREADS - 1) / CTA_THREADS), CTA_THREADS, 0, (cudaStream_t)stream>>>(
        e->store->d, e->s, n, mask_dev, flags_dev, region_map_dev, region_fit_dev);
    CU(cudaGetLastError());
    e->launches++;
    return SPARC_OK;
}

int sparc_stats(SparcEnv* e, int64_t* out_dev, int clear, void* stream) {
    if (!e || !out_dev) return fail(SPARC_E_ARG, "null argument%s");
    DeviceGuard dg_(e->store->device); CU(dg_.err);
    k_read_stats<<<1, 32 * SPARC_N_STATS, 0, (cudaStream_t)stream>>>(e->s.stats, (long long*)out_dev, clear);
    CU(cudaGetLastError());
    e->launches++;
    return SPARC_OK;
}

// device address of pinned (page-locked) host memory under unified addressing, or nullptr for pageable memory
static void* mapped_ptr(const void* p) {
    cudaPointerAttributes a;
    if (!p || cudaPointerGetAttributes(&a, p) != cudaSuccess || a.type != cudaMemoryTypeHost) { cudaGetLastError(); return nullptr; }
    return a.devicePointer;
}

#ifdef HOST_TIMING
#include <time.h>
static double g_ht[5];
static inline double ht_now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3; }
#define HT_T(x) const double x = ht_now()
#define HT_ACC() do { g_ht[0] += t1 - t0; g_ht[1] += t2 - t1; g_ht[2] += t3 - t2; g_ht[3] += t4 - t3; g_ht[4] += 1; } while (0)
extern "C" void sparc_debug_hosttiming(double* out5) { for (int k = 0; k < 5; k++) { out5[k] = g_ht[k]; g_ht[k] = 0; } }
#else
#define HT_T(x) ((void)0)
#define HT_ACC() ((void)0)
#endif
int sparc_step_host(SparcEnv* e, const uint8_t* actions_host, float* reward_host, uint8_t* terminated_host,
                    uint8_t* truncated_host, const SparcObsOut* obs, void* stream) {
    NvtxRange nvtx_("sparc_step_host");
    if (!e || !actions_host) return fail(SPARC_E_ARG, "null argument%s");
    DeviceGuard dg_(e->store->device); CU(dg_.err);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)e->cfg.n_envs;
    SparcStepOut out; memset(&out, 0, sizeof(out));
    const char* ze = getenv("SPARC_HOST_ZERO_COPY");             // 0: the plain copy sequence below (A/B)
    if (e->split_chunks == 0 && e->res_host && !(ze && atoi(ze) == 0)) {
        // Launch-bound batches (the fused kernel): ONE launch, no copies, no synchronise.  The actions ride in the
        // launch itself (two bits each in the kernel parameters: no PCIe read in front of the dynamics), every env's
        // result is one byte in device memory that the last CTA to finish its dynamics expands into reward / terminated
        // / truncated arrays in mapped host memory (one coalesced sweep) before it raises a flag there, and the host
        // spins on the flag and copies the arrays out.  The call returns when the results are in the
        // caller's buffers (pageable or pinned alike); the observation rewrite of this step, which stays on the device,
        // overlaps the caller's turn-around and is ordered before anything put on `stream` later.
        // Measured at 4,096 envs (config 2), us per call: four copies around the launch 44.7; the kernel reading and
        // writing the caller's pinned buffers over PCIe (one sector per CTA and array) + cudaStreamSynchronize 31.5 --
        // the same with a completion flag instead of the synchronise 30.4: the ~4,000 small PCIe transactions were
        // the cost, not the wake-up.
        bool ok = true;
        HT_T(t0);
        const size_t nw = (n + 7) / 8;
        for (size_t k = 0; k < nw; k++) {                        // eight actions -> sixteen bits
            uint64_t x = 0;
            const size_t left = n - 8 * k;
            memcpy(&x, actions_host + 8 * k, left < 8 ? left : 8);
            if (x & 0xFCFCFCFCFCFCFCFCull) { ok = false; break; }    // an action outside 0..3 (an illegal move): plain path
            x = (x | (x >> 6)) & 0x000F000F000F000Full;
            x = (x | (x >> 12)) & 0x000000FF000000FFull;
            x = (x | (x >> 24)) & 0xFFFFull;
            uint32_t& w = e->pa.w[k >> 1];
            w = (k & 1) ? (w | ((uint32_t)x << 16)) : (uint32_t)x;
        }
        if (ok) {
            e->notify = true;
            e->done_seq += 1;
            HT_T(t1);
            int rc = sparc_step(e, e->act_stage, SPARC_ACT_U8, &out, obs, stream);    // (the action pointer is not read)
            e->notify = false;
            if (rc) return rc;
            HT_T(t2);
            for (uint32_t spins = 1; *e->done_host != e->done_seq; spins++) {
                if ((spins & 0xFFFu) == 0 && cudaStreamQuery(s) != cudaErrorNotReady) break;   // finished or failed: settle below
#if defined(__x86_64__)
                __builtin_ia32_pause();
#endif
            }
            if (*e->done_host != e->done_seq) CU(cudaStreamSynchronize(s));                     // surfaces a launch / kernel error
            HT_T(t3);
            if (reward_host) memcpy(reward_host, e->res_host, n * 4);
            if (terminated_host) memcpy(terminated_host, e->res_host + 4 * e->res_nb, n);
            if (truncated_host) memcpy(truncated_host, e->res_host + 5 * e->res_nb, n);
            HT_T(t4);
            HT_ACC();
            return SPARC_OK;
        }
    }
    out.reward_dev = e->rew_stage; out.terminated_dev = e->term_stage; out.truncated_dev = e->trunc_stage;
    // Big batches: the result copies are issued right behind k_dyn on a side stream and overlap k_encode, and the call
    // returns when THEY are done -- not when the observation rewrite is, which stays on the device and is ordered before
    // anything put on `stream` later.  The caller's turn-around (and the next call's action upload, on the same side
    // stream) then overlaps the encoder instead of following it: the GPU never waits for the host.
    // SPARC_HOST_EARLY=0: synchronise the stream as before (A/B).
    const bool overlap = e->split_chunks == 1 && e->copy_stream;
    const char* ee = getenv("SPARC_HOST_EARLY");
    const bool early = overlap && !(ee && atoi(ee) == 0);
    if (early) {
        CU(cudaMemcpyAsync(e->act_stage, actions_host, n, cudaMemcpyHostToDevice, e->copy_stream));
        CU(cudaEventRecord(e->ev_act, e->copy_stream));
        CU(cudaStreamWaitEvent(s, e->ev_act, 0));
    } else {
        CU(cudaMemcpyAsync(e->act_stage, actions_host, n, cudaMemcpyHostToDevice, s));
    }
    if (overlap) { e->hostcopy.reward = reward_host; e->hostcopy.term = terminated_host; e->hostcopy.trunc = truncated_host; e->hostcopy.active = true; }
    int rc = sparc_step(e, e->act_stage, SPARC_ACT_U8, &out, obs, stream);
    e->hostcopy.active = false;
    if (rc) return rc;
    if (!overlap) {
        if (reward_host) CU(cudaMemcpyAsync(reward_host, e->rew_stage, n * 4, cudaMemcpyDeviceToHost, s));
        if (terminated_host) CU(cudaMemcpyAsync(terminated_host, e->term_stage, n, cudaMemcpyDeviceToHost, s));
        if (truncated_host) CU(cudaMemcpyAsync(truncated_host, e->trunc_stage, n, cudaMemcpyDeviceToHost, s));
    }
    if (early) CU(cudaEventSynchronize(e->ev_copy));
    else CU(cudaStreamSynchronize(s));
    return SPARC_OK;
}

static void* mapped_ptr(const void* p);
int64_t sparc_info_record_bytes(const SparcEnv* e) {
    if (!e) return 0;
    return e->cfg.n_envs * (int64_t)(4 + 2 + 2 + 4 + SPARC_MAX_REGIONS + e->store->d.S);
}

static int info_call(SparcEnv* e, bool reset, const int32_t* ints_host, void* record_host, const SparcObsOut* obs, void* stream) {
    NvtxRange nvtx_("sparc_info_host");
    if (!e || !record_host || (!reset && !ints_host)) return fail(SPARC_E_ARG, "null argument%s");
    DeviceGuard dg_(e->store->device); CU(dg_.err);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)e->cfg.n_envs, bytes = (size_t)sparc_info_record_bytes(e);
    if (n <= SPARC_INFO_MAX_ENVS && e->done_host && obs &&
        (obs->format == SPARC_OBS_REFERENCE || obs->format == SPARC_OBS_COMPACT)) {
        // a handful of envs (the single-env drop-in): one launch of k_info, the arguments in the launch, the record
        // written into mapped host memory by the kernel, a completion flag instead of a synchronise (see k_info)
        if (!e->info_host) {
            CU(cudaHostAlloc((void**)&e->info_host, bytes, cudaHostAllocMapped));
            CU(cudaHostGetDevicePointer((void**)&e->info_host_dev, (void*)e->info_host, 0));
        }
        KParams p; memset(&p, 0, sizeof(p));
        p.st = e->store->d; p.sd = e->s; p.cfg = e->cfg; p.tile = 1;
        p.mode = reset ? 1 : 0;
        p.obs = *obs; p.obs.update = SPARC_OBS_REWRITE;
        SmallInts args;
        for (size_t k = 0; k < n; k++) args.v[k] = ints_host ? (long long)ints_host[k] : SPARC_NO_ARG;
        e->done_seq += 1;
        p.done_flag = e->done_dev; p.done_count = e->done_count; p.done_seq = e->done_seq;
        const size_t shmem = (size_t)(tile_head_words(1) + tile_flat_words(p.st, 1)) * sizeof(uint32_t) + SPARC_MAX_REGIONS + (size_t)p.st.S;
        if (obs->format == SPARC_OBS_COMPACT) k_info<SPARC_OBS_COMPACT><<<(unsigned)n, CTA_THREADS, shmem, s>>>(p, args, e->info_host_dev);
        else k_info<SPARC_OBS_REFERENCE><<<(unsigned)n, CTA_THREADS, shmem, s>>>(p, args, e->info_host_dev);
        CU(cudaGetLastError());
        e->launches++;
        for (uint32_t spins = 1; *e->done_host != e->done_seq; spins++) {
            if ((spins & 0xFFFu) == 0 && cudaStreamQuery(s) != cudaErrorNotReady) break;       // finished or failed: settle below
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
        if (*e->done_host != e->done_seq) CU(cudaStreamSynchronize(s));                         // surfaces a launch / kernel error
        memcpy(record_host, e->info_host, bytes);
        return SPARC_OK;
    }
    // more envs: pinned host buffers are read and written by the kernels themselves (a few hundred bytes per env over
    // PCIe), pageable ones go through device staging
    char* b = (char*)mapped_ptr(record_host);
    const int32_t* ints = ints_host ? (const int32_t*)mapped_ptr(ints_host) : nullptr;
    const bool direct = b && (!ints_host || ints) && e->split_chunks == 0;
    if (!direct) {
        if (!e->info_stage) {
            CU(cudaMalloc((void**)&e->info_stage, bytes));
            CU(cudaMalloc((void**)&e->int_stage, n * 4));
        }
        b = e->info_stage;
        if (ints_host) CU(cudaMemcpyAsync(e->int_stage, ints_host, n * 4, cudaMemcpyHostToDevice, s));
        ints = ints_host ? e->int_stage : nullptr;
    }
    SparcStepOut out; memset(&out, 0, sizeof(out));
    out.step_count_dev = (int32_t*)b;
    uint16_t* flags = (uint16_t*)(b + 4 * n);
    out.agent_xy_dev = (uint8_t*)(b + 6 * n);
    out.reward_code_dev = (int8_t*)(b + 8 * n);
    out.terminated_dev = (uint8_t*)(b + 9 * n);
    out.truncated_dev = (uint8_t*)(b + 10 * n);
    out.legal_mask_dev = (uint8_t*)(b + 11 * n);
    int8_t* fit = (int8_t*)(b + 12 * n);
    int8_t* map = fit + (size_t)SPARC_MAX_REGIONS * n;
    int rc = reset ? sparc_reset(e, ints, nullptr, &out, obs, stream) : sparc_step(e, ints, SPARC_ACT_I32, &out, obs, stream);
    if (rc) return rc;
    if ((rc = sparc_validate(e, nullptr, flags, map, fit, stream))) return rc;
    if (!direct) CU(cudaMemcpyAsync(record_host, b, bytes, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return SPARC_OK;
}
int sparc_step_info_host(SparcEnv* e, const int32_t* actions_host, void* record_host, const SparcObsOut* obs, void* stream) {
    return info_call(e, false, actions_host, record_host, obs, stream);
}
int sparc_reset_info_host(SparcEnv* e, const int32_t* puzzle_idx_host, void* record_host, const SparcObsOut* obs, void* stream) {
    return info_call(e, true, puzzle_idx_host, record_host, obs, stream);
}

int64_t sparc_state_bytes(const SparcEnv* e) {
    if (!e) return 0;
    const int64_t n = e->cfg.n_envs;
    return n * (16 + 32 + 4 + (e->s.moves ? SPARC_MOVE_WORDS * 4 : 0));
}

static int state_copy(SparcEnv* e, char* blob, bool save, void* stream) {
    DeviceGuard dg_(e->store->device); CU(dg_.err);
    const size_t n = (size_t)e->cfg.n_envs;
    struct { void* p; size_t bytes; } parts[] = {{e->s.core, n * 16}, {e->s.vis, n * 32}, {e->s.ep_ret, n * 4},
                                                 {e->s.moves, e->s.moves ? n * SPARC_MOVE_WORDS * 4 : 0}};
    for (auto& part : parts) {
        if (!part.bytes) continue;
        CU(cudaMemcpyAsync(save ? (void*)blob : part.p, save ? part.p : (void*)blob, part.bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        blob += part.bytes;
    }
    return SPARC_OK;
}
int sparc_state_save(SparcEnv* e, void* blob_dev, void* stream) {
    if (!e || !blob_dev) return fail(SPARC_E_ARG, "null argument%s");
    return state_copy(e, (char*)blob_dev, true, stream);
}
int sparc_state_load(SparcEnv* e, const void* blob_dev, void* stream) {
    if (!e || !blob_dev) return fail(SPARC_E_ARG, "null argument%s");
    return state_copy(e, (char*)blob_dev, false, stream);
}

int sparc_debug_valprof(uint64_t* out8_host, int clear) {     /* -DVAL_PROFILE builds only: see VP_* in sparc_core.cuh, tools/valprof.py */
#ifdef VAL_PROFILE
    if (!out8_host) return fail(SPARC_E_ARG, "null argument%s");
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpyFromSymbol(out8_host, g_valprof, 8 * sizeof(uint64_t)));
    if (clear) { const unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0}; CU(cudaMemcpyToSymbol(g_valprof, z, sizeof(z))); }
    return SPARC_OK;
#else
    (void)out8_host; (void)clear;
    return fail(SPARC_E_ARG, "not a -DVAL_PROFILE build%s");
#endif
}

int sparc_debug_envprof(uint64_t* out16_host, int clear) {    /* -DENV_PROFILE builds only: see EP_* above, tools/envprof.py */
#ifdef ENV_PROFILE
    if (!out16_host) return fail(SPARC_E_ARG, "null argument%s");
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpyFromSymbol(out16_host, g_envprof, 16 * sizeof(uint64_t)));
    if (clear) { unsigned long long z[16] = {0}; z[11] = ~0ull; CU(cudaMemcpyToSymbol(g_envprof, z, sizeof(z))); }
    return SPARC_OK;
#else
    (void)out16_host; (void)clear;
    return fail(SPARC_E_ARG, "not a -DENV_PROFILE build%s");
#endif
}

int64_t sparc_env_num_envs(const SparcEnv* e) { return e ? e->cfg.n_envs : 0; }
int sparc_kernel_launches(const SparcEnv* e, int64_t* out_count) {
    if (!e || !out_count) return fail(SPARC_E_ARG, "null argument%s");
    *out_count = e->launches;
    return SPARC_OK;
}

}  // extern "C"
