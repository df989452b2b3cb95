// TEST-ONLY: compiles sparc_core.cuh (the product's device code) as plain C++ and replays the
// kernels' control flow serially, so the dynamics / validator / encoder logic can be checked
// against the oracle in a container without a GPU.  Nothing in sparc_gym_b200/ links or loads
// this; the product path is libsparc_b200.so (CUDA) only.
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../sparc_gym_b200/csrc/sparc_core.cuh"

static StoreDev make_store(const SparcStoreDesc* d) {
    StoreDev s;
    s.n_puzzles = d->n_puzzles; s.SX = d->slab_x; s.SY = d->slab_y; s.S = s.SX * s.SY;
    s.P = d->n_planes; s.L = s.P * s.S; s.W = (s.L + 31) / 32;
    s.Wpad = (s.W + 1 > 9 ? s.W + 1 : 9) | 1;
    s.nv = (s.S + 31) / 32;
    s.hdr = d->headers; s.nb = d->node_boards; s.cb = d->cell_boards; s.sbits = d->static_bits;
    s.color = d->color; s.info = d->info; s.shapes = d->shapes; s.trie = d->trie;
    return s;
}

extern "C" {

// state arrays are host memory laid out exactly like the device SoA
int emu_launch(const SparcStoreDesc* desc, uint32_t* core, uint32_t* vis, int32_t* ep_ret, uint32_t* moves,
               unsigned long long* stats, long long n_envs, int traceback, int max_steps, int autoreset,
               int stride, int validate_mode, int mode, int tile, int nthr, const void* actions, int act_dtype,
               const int32_t* reset_idx, const uint8_t* reset_mask, const SparcStepOut* out, const SparcObsOut* obs) {
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
    const StoreDev& st = p.st;
    std::vector<uint32_t> smem((size_t)tile * st.Wpad + tile + 16, 0xDEADBEEFu);
    std::vector<EnvRegs> regs(tile);
    for (long long env0 = 0; env0 < n_envs; env0 += tile) {
        const int nE = (int)((n_envs - env0) < tile ? (n_envs - env0) : tile);
        uint32_t* rows = smem.data();
        uint32_t* tile_puzzle = rows + (size_t)tile * st.Wpad;
        for (int tid = 0; tid < nE; tid++) {
            unsigned long long my[SPARC_N_STATS];
            if (phase_a(p, env0 + tid, regs[tid], rows + tid * st.Wpad, my) && mode == 0)
                for (int k = 0; k < SPARC_N_STATS; k++) stats[k] += my[k];
            tile_puzzle[tid] = regs[tid].puzzle;
        }
        if (p.obs.base_dev) {
            for (int tid = 0; tid < nthr; tid++) stage_static(st, rows, tile_puzzle, nE, tid, nthr);
            for (int tid = 0; tid < nE; tid++) stage_dynamic(st, rows, tid, regs[tid]);
            for (int tid = 0; tid < nthr; tid++) expand_base(st, rows, p.obs.base_dev + (size_t)env0 * st.L, nE, tid, nthr);
        }
        if (p.obs.color_dev) for (int tid = 0; tid < nthr; tid++) expand_color(st, tile_puzzle, p.obs.color_dev + (size_t)env0 * st.S, nE, tid, nthr);
        if (p.obs.info_dev) for (int tid = 0; tid < nthr; tid++) expand_info(st, tile_puzzle, p.obs.info_dev + (size_t)env0 * st.S, nE, tid, nthr);
    }
    return 0;
}

int emu_validate(const SparcStoreDesc* desc, uint32_t* core, uint32_t* vis, int32_t* ep_ret, long long n_envs,
                 uint16_t* flags, int8_t* region_map) {
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
        flags[i] = (uint16_t)(validate_env(st, e.puzzle, vd, e.x, e.y, rm) | SPARC_RULE_VALID);
    }
    return 0;
}

}  // extern "C"
