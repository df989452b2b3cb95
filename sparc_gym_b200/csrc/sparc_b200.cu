// sparc_b200.cu -- sm_100a kernels + the C ABI of include/sparc_b200.h.
//
// One fused kernel per env step: a CTA owns a tile of consecutive envs; the first threads run
// the dynamics (one env per thread: legality, move / traceback undo, trie-cursor reward,
// truncation, terminal rule validation, statistics, autoreset), then the whole CTA expands the
// tile's bit-packed planes into the caller's int32/int64 observation tensors with coalesced
// 16-byte streaming stores.  The kernel is HBM-write bound: per env-step it
// writes X*Y*(4P+4+8) observation bytes and touches ~100 bytes of state.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <new>

#include "sparc_core.cuh"

#define CTA_THREADS 128
#ifndef SPARC_MINBLOCKS          // resident CTAs per SM the register allocator must allow.  8 -> 64
#define SPARC_MINBLOCKS 8        // registers: the rarely-run validator spills, the streaming encoder
#endif                           // gets 2x the resident warps (measured: 0.322 -> 0.219 ms/step, config 4)

static thread_local char g_err[512] = "";
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
    StoreDev d;
    void* blocks[9];
    int plane_ids[SPARC_MAX_PLANES];
};

struct SparcEnv {
    SparcStore* store;
    StepCfg cfg;
    StateDev s;
    int tile;                // envs per CTA
    int64_t launches;
    uint8_t* act_stage;      // device staging for the *_host call
    float* rew_stage;
    uint8_t* term_stage;
    uint8_t* trunc_stage;
};

// dynamic shared (tile_smem_words): dyn[tile][9] | tile_puzzle[tile] | flat | cflat | zflat
extern __shared__ uint32_t smem[];

__global__ void __launch_bounds__(CTA_THREADS, SPARC_MINBLOCKS) k_env(const KParams p) {
    const StoreDev& st = p.st;
    const int tid = threadIdx.x;
    const long long env0 = (long long)blockIdx.x * p.tile;
    const long long n = p.cfg.n_envs;
    const int nE = (int)min((long long)p.tile, n - env0);
    uint32_t* dyn = smem;
    uint32_t* tile_puzzle = smem + p.tile * SPARC_DYN_STRIDE;

    EnvRegs e;
    unsigned long long my_stats[SPARC_N_STATS];
    bool have_stats = false;
    if (tid < nE) {
        have_stats = phase_a(p, env0 + tid, e, dyn + tid * SPARC_DYN_STRIDE, my_stats);
        tile_puzzle[tid] = e.puzzle;
        write_dyn(st, dyn, tid, e);
    }
    // statistics: warp-reduce, one atomic per counter per warp (the dynamics threads are the first warps)
    if (p.mode == 0 && tid < ((nE + 31) & ~31)) {
#pragma unroll
        for (int k = 0; k < SPARC_N_STATS; k++) {
            unsigned long long v = have_stats ? my_stats[k] : 0ull;
#pragma unroll
            for (int off = 16; off; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
            if ((tid & 31) == 0 && v) atomicAdd(p.sd.stats + k, v);
        }
    }
    if (!p.obs.base_dev && !p.obs.color_dev && !p.obs.info_dev) return;
    __syncthreads();
    encode_tile_stage1(st, smem, p.tile, nE, p.obs, tid, CTA_THREADS);
    __syncthreads();
    encode_tile_stage2(st, smem, p.tile, nE, env0, p.obs, tid, CTA_THREADS);
}

// stand-alone validator (SG.py:941-950) for sparc_validate: one env per thread
__global__ void __launch_bounds__(CTA_THREADS) k_validate(const StoreDev st, const StateDev sd, long long n,
                                                          const uint8_t* mask, uint16_t* flags, int8_t* region_map) {
    const long long i = (long long)blockIdx.x * CTA_THREADS + threadIdx.x;
    if (i >= n) return;
    if (mask && !mask[i]) { if (flags) flags[i] = 0; return; }
    EnvRegs e;
    load_env(e, sd, st, n, i);
    uint32_t vd[9];
#pragma unroll
    for (int k = 0; k < 8; k++) vd[k] = e.v[k];
    vd[8] = 0;
    int8_t* rm = region_map ? region_map + (size_t)i * st.S : nullptr;
    if (rm) for (int k = 0; k < st.S; k++) rm[k] = -1;
    const int f = validate_env(st, e.puzzle, vd, e.x, e.y, rm);
    if (flags) flags[i] = (uint16_t)(f | SPARC_RULE_VALID);
}

__global__ void k_init_state(StateDev sd, long long n, int n_puzzles) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    sd.core[4 * i + 0] = (uint32_t)(i % n_puzzles);
    sd.core[4 * i + 1] = 0; sd.core[4 * i + 2] = 0; sd.core[4 * i + 3] = 0;
}

__global__ void k_read_stats(unsigned long long* stats, long long* out, int clear) {
    const int k = threadIdx.x;
    if (k < SPARC_N_STATS) { out[k] = (long long)stats[k]; if (clear) stats[k] = 0; }
}

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
    CU(cudaSetDevice(device));
    SparcStore* s = new (std::nothrow) SparcStore();
    if (!s) return fail(SPARC_E_NOMEM, "out of host memory%s");
    memset(s, 0, sizeof(*s));
    s->device = device;
    StoreDev& d = s->d;
    d.n_puzzles = desc->n_puzzles; d.SX = desc->slab_x; d.SY = desc->slab_y; d.S = d.SX * d.SY;
    d.P = desc->n_planes; d.L = d.P * d.S; d.W = (d.L + 31) / 32;
    if (desc->words_per_puzzle != d.W) { delete s; return fail(SPARC_E_ARG, "words_per_puzzle mismatch%s"); }
    d.nv = (d.S + 31) / 32;
    d.inv_L = (uint32_t)((1ull << 32) / (uint64_t)d.L) + 1u;
    d.inv_S = (uint32_t)((1ull << 32) / (uint64_t)d.S) + 1u;
    d.inv_4S = (uint32_t)((1ull << 32) / (uint64_t)(4 * d.S)) + 1u;
    d.Wc = (4 * d.S + 31) / 32;
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
    *out = s;
    return SPARC_OK;
}

int sparc_store_destroy(SparcStore* s) {
    if (!s) return SPARC_OK;
    cudaSetDevice(s->device);
    for (void* b : s->blocks) if (b) cudaFree(b);
    delete s;
    return SPARC_OK;
}

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
    if (autoreset_mode < 0 || autoreset_mode > 2 || validate_mode < 0 || validate_mode > 2 || puzzle_stride < 0)
        return fail(SPARC_E_ARG, "bad mode%s");
    CU(cudaSetDevice(store->device));
    SparcEnv* e = new (std::nothrow) SparcEnv();
    if (!e) return fail(SPARC_E_NOMEM, "out of host memory%s");
    memset(e, 0, sizeof(*e));
    e->store = store;
    e->cfg.n_envs = n_envs; e->cfg.traceback = traceback != 0; e->cfg.max_steps = max_steps;
    e->cfg.autoreset = autoreset_mode; e->cfg.stride = puzzle_stride; e->cfg.validate_mode = validate_mode;
    e->tile = pick_tile(n_envs);
    const size_t n = (size_t)n_envs;
    cudaError_t err = cudaSuccess;
    auto alloc = [&](void** p, size_t bytes) { if (err == cudaSuccess) { err = cudaMalloc(p, bytes); if (err == cudaSuccess) err = cudaMemset(*p, 0, bytes); } };
    alloc((void**)&e->s.core, n * 16);
    alloc((void**)&e->s.vis, n * 8 * 4);
    alloc((void**)&e->s.ep_ret, n * 4);
    if (traceback) alloc((void**)&e->s.moves, n * SPARC_MOVE_WORDS * 4);
    alloc((void**)&e->s.stats, SPARC_N_STATS * 8);
    alloc((void**)&e->act_stage, n);
    alloc((void**)&e->rew_stage, n * 4);
    alloc((void**)&e->term_stage, n);
    alloc((void**)&e->trunc_stage, n);
    if (err != cudaSuccess) { sparc_env_destroy(e); return fail(SPARC_E_CUDA, "cudaMalloc: %s", cudaGetErrorString(err)); }
    k_init_state<<<(unsigned)((n + 255) / 256), 256>>>(e->s, n_envs, store->d.n_puzzles);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    *out = e;
    return SPARC_OK;
}

int sparc_env_destroy(SparcEnv* e) {
    if (!e) return SPARC_OK;
    cudaSetDevice(e->store->device);
    void* ptrs[] = {e->s.core, e->s.vis, e->s.ep_ret, e->s.moves, e->s.stats, e->act_stage, e->rew_stage, e->term_stage, e->trunc_stage};
    for (void* p : ptrs) if (p) cudaFree(p);
    delete e;
    return SPARC_OK;
}

static int launch_env(SparcEnv* e, KParams& p, void* stream) {
    p.st = e->store->d; p.sd = e->s; p.cfg = e->cfg; p.tile = e->tile;
    const int64_t n = e->cfg.n_envs;
    const unsigned grid = (unsigned)((n + e->tile - 1) / e->tile);
    const size_t shmem = (size_t)tile_smem_words(p.st, e->tile) * sizeof(uint32_t);
    k_env<<<grid, CTA_THREADS, shmem, (cudaStream_t)stream>>>(p);
    CU(cudaGetLastError());
    e->launches++;
    return SPARC_OK;
}

int sparc_reset(SparcEnv* e, const int32_t* puzzle_idx_dev, const uint8_t* mask_dev, const SparcStepOut* info_out,
                const SparcObsOut* obs, void* stream) {
    if (!e) return fail(SPARC_E_ARG, "null env%s");
    KParams p; memset(&p, 0, sizeof(p));
    p.mode = 1; p.reset_idx = puzzle_idx_dev; p.reset_mask = mask_dev;
    if (info_out) p.out = *info_out;
    if (obs) p.obs = *obs;
    return launch_env(e, p, stream);
}

int sparc_step(SparcEnv* e, const void* actions_dev, int action_dtype, const SparcStepOut* out, const SparcObsOut* obs, void* stream) {
    if (!e || !actions_dev) return fail(SPARC_E_ARG, "null argument%s");
    if (action_dtype < 0 || action_dtype > 2) return fail(SPARC_E_ARG, "bad action dtype%s");
    KParams p; memset(&p, 0, sizeof(p));
    p.mode = 0; p.actions = actions_dev; p.act_dtype = action_dtype;
    if (out) p.out = *out;
    if (obs) p.obs = *obs;
    return launch_env(e, p, stream);
}

int sparc_validate(SparcEnv* e, const uint8_t* mask_dev, uint16_t* flags_dev, int8_t* region_map_dev, void* stream) {
    if (!e) return fail(SPARC_E_ARG, "null env%s");
    const int64_t n = e->cfg.n_envs;
    k_validate<<<(unsigned)((n + CTA_THREADS - 1) / CTA_THREADS), CTA_THREADS, 0, (cudaStream_t)stream>>>(
        e->store->d, e->s, n, mask_dev, flags_dev, region_map_dev);
    CU(cudaGetLastError());
    e->launches++;
    return SPARC_OK;
}

int sparc_stats(SparcEnv* e, int64_t* out_dev, int clear, void* stream) {
    if (!e || !out_dev) return fail(SPARC_E_ARG, "null argument%s");
    k_read_stats<<<1, 32, 0, (cudaStream_t)stream>>>(e->s.stats, (long long*)out_dev, clear);
    CU(cudaGetLastError());
    e->launches++;
    return SPARC_OK;
}

int sparc_step_host(SparcEnv* e, const uint8_t* actions_host, float* reward_host, uint8_t* terminated_host,
                    uint8_t* truncated_host, const SparcObsOut* obs, void* stream) {
    if (!e || !actions_host) return fail(SPARC_E_ARG, "null argument%s");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)e->cfg.n_envs;
    CU(cudaMemcpyAsync(e->act_stage, actions_host, n, cudaMemcpyHostToDevice, s));
    SparcStepOut out; memset(&out, 0, sizeof(out));
    out.reward_dev = e->rew_stage; out.terminated_dev = e->term_stage; out.truncated_dev = e->trunc_stage;
    int rc = sparc_step(e, e->act_stage, SPARC_ACT_U8, &out, obs, stream);
    if (rc) return rc;
    if (reward_host) CU(cudaMemcpyAsync(reward_host, e->rew_stage, n * 4, cudaMemcpyDeviceToHost, s));
    if (terminated_host) CU(cudaMemcpyAsync(terminated_host, e->term_stage, n, cudaMemcpyDeviceToHost, s));
    if (truncated_host) CU(cudaMemcpyAsync(truncated_host, e->trunc_stage, n, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return SPARC_OK;
}

int64_t sparc_state_bytes(const SparcEnv* e) {
    if (!e) return 0;
    const int64_t n = e->cfg.n_envs;
    return n * (16 + 32 + 4 + (e->s.moves ? SPARC_MOVE_WORDS * 4 : 0));
}

static int state_copy(SparcEnv* e, char* blob, bool save, void* stream) {
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

int64_t sparc_env_num_envs(const SparcEnv* e) { return e ? e->cfg.n_envs : 0; }
int sparc_kernel_launches(const SparcEnv* e, int64_t* out_count) {
    if (!e || !out_count) return fail(SPARC_E_ARG, "null argument%s");
    *out_count = e->launches;
    return SPARC_OK;
}

}  // extern "C"
