// TEST-ONLY: compiles sparc_core.cuh (the product's device code) as plain C++ and replays the
// kernels' control flow serially, so the dynamics / validator / encoder logic can be checked
// against the oracle in a container without a GPU.  Nothing in sparc_gym_b200/ links or loads
// this; the product path is libsparc_b200.so (CUDA) only.
#include <stdlib.h>
#include <string.h>
#include <vector>
#define SPARC_EMU_COUNTERS 1
#include "../../sparc_gym_b200/csrc/sparc_core.cuh"

static StoreDev make_store(const SparcStoreDesc* d) {
    StoreDev s;
    s.n_puzzles = d->n_puzzles; s.SX = d->slab_x; s.SY = d->slab_y; s.S = s.SX * s.SY;
    s.P = d->n_planes; s.L = s.P * s.S; s.W = (s.L + 31) / 32;
    s.nv = (s.S + 31) / 32;
    s.inv_L = (uint32_t)((1ull << 32) / (uint64_t)s.L) + 1u;
    s.inv_S = (uint32_t)((1ull << 32) / (uint64_t)s.S) + 1u;
    s.inv_4S = (uint32_t)((1ull << 32) / (uint64_t)(4 * s.S)) + 1u;
    s.Wc = (4 * s.S + 31) / 32;
    { const uint64_t k0b = (uint64_t)(s.L >> 5), k0c = (uint64_t)((4 * s.S) >> 5), k0z = (uint64_t)(s.S >> 5);
      s.inv_k0b = k0b ? (uint32_t)((1ull << 32) / k0b) + 1u : 0u; s.inv_k0c = k0c ? (uint32_t)((1ull << 32) / k0c) + 1u : 0u;
      s.inv_k0z = k0z ? (uint32_t)((1ull << 32) / k0z) + 1u : 0u; }
    s.hdr = d->headers; s.nb = d->node_boards; s.cb = d->cell_boards; s.sbits = d->static_bits;
    s.color4 = d->color4; s.info_nz = d->info_nz; s.info = d->info; s.shapes = d->shapes; s.trie = d->trie; s.vrec = nullptr;
    return s;
}

extern "C" {

// state arrays are host memory laid out exactly like the device SoA
int emu_launch(const SparcStoreDesc* desc, uint32_t* core, uint32_t* vis, int32_t* ep_ret, uint32_t* moves,
               unsigned long long* stats, long long n_envs, int traceback, int max_steps, int autoreset,
               int stride, int validate_mode, int mode, int tile, int nthr, const void* actions, int act_dtype,
               const int32_t* reset_idx, const uint8_t* reset_mask, const SparcStepOut* out, const SparcObsOut* obs,
               int defer_validation) {
    KParams p;
    memset(&p, 0, sizeof(p));
    p.st = make_store(desc);
    p.sd.core = core; p.sd.vis = vis; p.sd.ep_ret = ep_ret; p.sd.moves = moves; p.sd.stats = stats;
    p.cfg.n_envs = n_envs; p.cfg.traceback = traceback; p.cfg.max_steps = max_steps; p.cfg.autoreset = autoreset;
    p.cfg.stride = stride; p.cfg.validate_mode = validate_mode;
    p.mode = mode; p.tile = tile; p.actions = actions; p.act_dtype = act_dtype;
    p.reset_idx = reset_idx; p.reset_mask = reset_mask;
    if (out) p.out = *out;
    if (obs) p.obs = *obs;
    std::vector<uint32_t> vq;
    uint32_t vq_count = 0;
    if (defer_validation) { vq.resize((size_t)n_envs * SPARC_VQ_WORDS); p.vq = vq.data(); p.vq_count = &vq_count; }
    // update = SPARC_OBS_DELTA: with defer_validation (the split step: k_dyn, then k_val || k_encode_list) the envs that
    // began a new episode are queued and rewritten one by one; otherwise (the fused k_env) a tile is rewritten when
    // one of its envs did
    const bool delta = mode == 0 && p.obs.update == SPARC_OBS_DELTA;
    if (mode == 1) p.obs.update = SPARC_OBS_REWRITE;
    std::vector<uint32_t> rq;
    uint32_t rq_count = 0;
    if (delta && defer_validation) { rq.resize((size_t)n_envs); p.rq = rq.data(); p.rq_count = &rq_count; }
    const StoreDev& st = p.st;
    // exactly the dynamic shared memory the kernels request: an out-of-bounds row access is an ASan error
    std::vector<uint32_t> smem((size_t)tile_head_words(tile) + tile_flat_words(st, tile) + tile_snap_words(tile), 0xDEADBEEFu);
    std::vector<EnvRegs> regs(tile);
    for (long long env0 = 0; env0 < n_envs; env0 += tile) {
        const int nE = (int)((n_envs - env0) < tile ? (n_envs - env0) : tile);
        uint32_t* dyn = smem.data();
        uint32_t* tile_puzzle = dyn + (size_t)tile * SPARC_DYN_STRIDE;
        uint32_t* tflags = dyn + tile_head_words(tile) + tile_flat_words(st, tile);
        uint32_t* snaps = tflags + tile;
        bool any_full = false;
        for (int tid = 0; tid < nE; tid++) {
            unsigned long long my[SPARC_N_STATS];
            int fl = 0;
            // the fused k_env (no global queue) leaves the states to validate in shared memory for its validator warp
            const bool have = defer_validation ? phase_a(p, env0 + tid, regs[tid], dyn + tid * SPARC_DYN_STRIDE, my, &fl)
                                               : phase_a<false, true>(p, env0 + tid, regs[tid], dyn + tid * SPARC_DYN_STRIDE, my, &fl,
                                                                      snaps + tid * SPARC_VQ_WORDS);
            if (have && mode == 0)
                for (int k = 0; k < SPARC_N_STATS; k++) stats[k] += my[k];
            tflags[tid] = (uint32_t)fl;
            any_full |= (fl & PA_NEW_EPISODE) != 0;
            tile_puzzle[tid] = regs[tid].puzzle;
            write_dyn(st, dyn, tid, regs[tid]);
        }
        for (int j = 0; j < nE; j++)
            if (tflags[j] & PA_SNAPSHOT) {
                const uint32_t* q = snaps + j * SPARC_VQ_WORDS;
                stats[SPARC_STAT_RULES_PASS] += (unsigned long long)(validate_snapshot(p, q) & (int)q[3]);
            }
        if (delta && (defer_validation || !any_full)) continue;
        uint32_t* flats = dyn + tile_head_words(tile);
        if (p.obs.format == SPARC_OBS_COMPACT) {
            for (int tid = 0; tid < nthr; tid++) encode_tile_stage1<SPARC_OBS_COMPACT>(st, dyn, flats, tile, nE, p.obs, tid, nthr);
            for (int tid = 0; tid < nthr; tid++) encode_tile_stage2<SPARC_OBS_COMPACT>(st, dyn, flats, tile, nE, env0, p.obs, tid, nthr);
        } else {
            for (int tid = 0; tid < nthr; tid++) encode_tile_stage1<SPARC_OBS_REFERENCE>(st, dyn, flats, tile, nE, p.obs, tid, nthr);
            for (int tid = 0; tid < nthr; tid++) encode_tile_stage2<SPARC_OBS_REFERENCE>(st, dyn, flats, tile, nE, env0, p.obs, tid, nthr);
        }
    }
    for (uint32_t slot = 0; slot < vq_count; slot++) stats[SPARC_STAT_RULES_PASS] += validate_queued(p, slot);
    if (delta && defer_validation) {                                  // k_encode_list
        std::vector<uint32_t> sm1((size_t)tile_head_words(1) + tile_flat_words(st, 1), 0xDEADBEEFu);
        uint32_t* head = sm1.data();
        uint32_t* flats = head + tile_head_words(1);
        for (uint32_t slot = 0; slot < rq_count; slot++) {
            const long long env = rq[slot];
            EnvRegs e;
            load_env(e, p.sd, st, n_envs, env);
            head[SPARC_DYN_STRIDE] = e.puzzle;
            write_dyn(st, head, 0, e);
            for (int tid = 0; tid < nthr; tid++) encode_env_stage1(st, head, flats, p.obs, tid, nthr);
            for (int tid = 0; tid < nthr; tid++) {
                if (p.obs.format == SPARC_OBS_COMPACT) encode_env_elementwise<SPARC_OBS_COMPACT>(st, head, flats, env, p.obs, tid, nthr);
                else encode_env_elementwise<SPARC_OBS_REFERENCE>(st, head, flats, env, p.obs, tid, nthr);
            }
        }
    }
    return 0;
}

int emu_validate(const SparcStoreDesc* desc, uint32_t* core, uint32_t* vis, int32_t* ep_ret, long long n_envs,
                 uint16_t* flags, int8_t* region_map, int8_t* region_fit) {
    StoreDev st = make_store(desc);
    StateDev sd; sd.core = core; sd.vis = vis; sd.ep_ret = ep_ret; sd.moves = nullptr; sd.stats = nullptr;
    for (long long i = 0; i < n_envs; i++) {
        EnvRegs e;
        load_env(e, sd, st, n_envs, i);
        uint32_t vd[9];
        for (int k = 0; k < 8; k++) vd[k] = e.v[k];
        vd[8] = 0;
        int8_t* rm = region_map ? region_map + (size_t)i * st.S : nullptr;
        if (rm) memset(rm, -1, st.S);
        int8_t* rf = region_fit ? region_fit + (size_t)i * SPARC_MAX_REGIONS : nullptr;
        if (rf) memset(rf, SPARC_FIT_NONE, SPARC_MAX_REGIONS);
        flags[i] = (uint16_t)(validate_env(st, e.puzzle, vd, e.x, e.y, rm, rf) | SPARC_RULE_VALID);
    }
    return 0;
}

// direct access to the exact-fit searches (which: 0 = sweep used by the kernels, 1 = reference order, 2 = sweep on wide counters)
int emu_exact_fit(int which, uint64_t region, int w, int h, const SparcShape* shapes, int ns,
                  const uint8_t* n_poly, const uint8_t* n_ylop, long long* iters) {
    g_fit_iters = 0;
    Cnt8 np = {{0, 0, 0, 0}}, ny = {{0, 0, 0, 0}};
    for (int s = 0; s < ns; s++) { cnt8_add(np, s, n_poly[s]); cnt8_add(ny, s, n_ylop[s]); }
    const bool ok = which == 2 ? exact_fit_budget(region, w, h, shapes, ns, np, ny, 0, true) != 0      // 6-bit demand counters
                  : which ? exact_fit_reference_order(region, w, h, shapes, ns, n_poly, n_ylop)
                          : exact_fit(region, w, h, shapes, ns, np, ny);
    if (iters) *iters = g_fit_iters;
    return ok ? 1 : 0;
}
// one copy of the split search (k_val runs the 2^nsplit codes on the lanes of a warp and ORs the answers)
int emu_exact_fit_split(uint64_t region, int w, int h, const SparcShape* shapes, int ns, const uint8_t* n_poly,
                        const uint8_t* n_ylop, int code, int nsplit, long long* iters) {
    g_fit_iters = 0;
    Cnt8 np = {{0, 0, 0, 0}}, ny = {{0, 0, 0, 0}};
    for (int s = 0; s < ns; s++) { cnt8_add(np, s, n_poly[s]); cnt8_add(ny, s, n_ylop[s]); }
    const bool ok = exact_fit(region, w, h, shapes, ns, np, ny, code, nsplit);
    if (iters) *iters = g_fit_iters;
    return ok ? 1 : 0;
}
void emu_fit_hist(long long* out24, int reset) {
    sparc_fit_book(); g_fit_cur = 0;
    for (int k = 0; k < 24; k++) { out24[k] = g_fit_hist[k]; if (reset) g_fit_hist[k] = 0; }
}
long long emu_fit_iters(int reset) { long long v = g_fit_iters; if (reset) g_fit_iters = 0; return v; }
void emu_fit_counters(long long* out3, int reset) {
    out3[0] = g_fit_iters; out3[1] = g_fit_calls; out3[2] = g_fit_max;
    if (reset) { g_fit_iters = g_fit_calls = g_fit_max = 0; }
}

}  // extern "C"
