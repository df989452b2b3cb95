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
//   * regions are flooded FOUR at a time (the four 8-lane groups start from seeds spread over the unlabelled cells);
//   * the exact polyomino fit is a converged-warp depth-first search with one shape slot per lane (exact_fit_coop: every
//     candidate placement of a node tested at once, two ballots pick the child); a search that outgrows its node budget
//     sends the snapshot to the hard queue, where k_val_hard runs VW_HARD_CODES shares of the search tree, one per warp (fit_step with
//     code = share, sparc_core.cuh) and the last share to finish combines the verdicts.
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

// One share of a HARD search (k_val): the whole warp runs copy `code` of the split search converged -- every lane
// the same fit_step (sparc_core.cuh) -- while the other copies run in other warps, anywhere on the chip.  `found` /
// `bit`: the item's found mask; a copy gives up as soon as another copy has found a fit for this region.
#ifndef VW_HARD_NSPLIT
#define VW_HARD_NSPLIT 5            /* 32 shares: measured p99 of the dense-board tail 3.0 / 2.0 / 1.7 / 2.0 / 2.9 / 4.6 ms with 8 / 16 / 32 / 64 / 128 / 256 shares */
#endif
#define VW_HARD_CODES (1 << VW_HARD_NSPLIT)
__device__ __noinline__ bool exact_fit_code(uint64_t region, int w, int h, const SparcShape* shapes, int ns,
                                            const Cnt8& n_poly, const Cnt8& n_ylop, int code,
                                            const unsigned long long* found, unsigned long long bit) {
    FitState f;
    int r = fit_init(f, region, w, h, shapes, ns, n_poly, n_ylop);
    for (int it = 0; r < 0; it++) {
        r = fit_step(f, shapes, ns, code, VW_HARD_NSPLIT);
        if ((it & 63) == 63 && (*reinterpret_cast<const volatile unsigned long long*>(found) & bit)) return true;
    }
    return r != 0;
}

__device__ __forceinline__ uint64_t vw_shfl64(uint64_t v, int src) {
    return vw_u64(__shfl_sync(VW_FULL, (uint32_t)v, src), __shfl_sync(VW_FULL, (uint32_t)(v >> 32), src));
}

// The same search as exact_fit() (sparc_core.cuh: one sweep over the cells, ylops and polys anchored AT the current
// cell in non-decreasing slot order), run by a converged warp: lane s owns shape slot s -- its mask / anchors and its
// instance counters live in that lane's registers -- so ALL candidate placements of a node are tested at once and two
// ballots pick the next child; the demand counters, the cursor and the totals are warp-uniform.  The DFS stack (cell,
// child) sits in shared memory, written by lane 0.  pa / yb: this lane's slot's poly / ylop instances inside the region.
// Returns 1 = fits, 0 = does not, -1 = node budget spent (the snapshot becomes a hard item of k_val).
__device__ __forceinline__ int exact_fit_coop(uint64_t region, int w, int h, const SparcShape* shapes, int ns,
                                              int pa, int yb, int lane, uint16_t* stack, int budget) {
    ShapeW sh = {0ull, 0ull};
    if (lane < ns) sh = load_shape(shapes + lane);
    const int fl = shp_flags(sh.mask);
    int rem_p = (fl & 2) ? 0 : pa, rem_y = (fl & 2) ? 0 : yb;                  // zero-area shapes never touch the grid
    if (__any_sync(VW_FULL, (rem_p | rem_y) != 0 && !(sh.anchors & SPARC_CELLS56))) return 0;   // fits nowhere on this board
    int tot_p = __reduce_add_sync(VW_FULL, rem_p), tot_y = __reduce_add_sync(VW_FULL, rem_y);
    if (tot_y > SPARC_MAX_YLOPS) return 0;                                     // (ingest guarantees this never happens)
    const bool all_connected = !__any_sync(VW_FULL, rem_p != 0 && !(fl & 4));
    const bool is_mono = (sh.mask & SPARC_CELLS56) == 1ull;
    const unsigned mono_mask = __ballot_sync(VW_FULL, is_mono);
    int tot_mono = __reduce_add_sync(VW_FULL, is_mono ? rem_p : 0);
    Cnt4 c = {region, 0, 0, 0};
    uint64_t grid = 0;
    for (int cx = 0; cx < w; cx++) grid |= ((1ull << h) - 1ull) << (cx * 8);
    int level = 0, cell = 0, q = 0;
    if (tot_y == 0) cell = __ffsll((long long)region) - 1;                     // no ylops: only demanded cells ever matter
    for (int nodes = 0; budget <= 0 || nodes < budget; nodes++) {
        const int n = cnt4_get(c, cell);
        const uint64_t later = grid & ~((2ull << cell) - 1ull);
        const uint64_t floor_cells = cnt4_eq(c, -tot_y) & later;
        const uint64_t open = (c.b0 | c.b1 | c.b2 | c.b3) & later;
        const uint64_t m = shape_at(sh, cell);
        const int q2 = q > ns ? q - ns : 0;
        const bool y_ok = rem_y > 0 && lane >= q && n + 1 <= tot_p && m != 0ull && !(floor_cells & ~m);
        bool p_ok = rem_p > 0 && lane >= q2 && n > 0 && m != 0ull && !(floor_cells & m);
        if (tot_y == 0 && all_connected) {             // isolated-cell prune (see fit_poly_ok)
            Cnt4 c2 = c;
            cnt4_sub(c2, m);
            const uint64_t D = (c2.b0 | c2.b1 | c2.b2) & ~c2.b3;
            const uint64_t iso = D & ~((D << 1) | (D >> 1) | (D << 8) | (D >> 8));
            if (iso && __popcll(iso) > tot_mono - (is_mono ? 1 : 0)) p_ok = false;
        }
        const unsigned vy = __ballot_sync(VW_FULL, y_ok), vp = __ballot_sync(VW_FULL, p_ok);
        if (vy | vp) {                                                         // first valid child: ylops before polys
            const bool is_y = vy != 0u;
            const int s = __ffs((int)(is_y ? vy : vp)) - 1;
            const uint64_t ms = vw_shfl64(m, s);
            if (is_y) { cnt4_add(c, ms); tot_y--; if (lane == s) rem_y--; }
            else { cnt4_sub(c, ms); tot_p--; if (lane == s) rem_p--; tot_mono -= (int)((mono_mask >> s) & 1u); }
            q = is_y ? s : ns + s;
            if (lane == 0) stack[level] = (uint16_t)(cell | (q << 8));
            level++;
            continue;
        }
        if (n == 0) {
            if (tot_p + tot_y == 0) { if (!open) return 1; }                   // nothing left and nothing demanded
            else if (later && (tot_y > 0 || open)) {                           // settled: advance (see fit_step)
                if (lane == 0) stack[level] = (uint16_t)(cell | (0xFF << 8));
                level++;
                cell = __ffsll((long long)(tot_y > 0 ? later : open)) - 1; q = 0;
                continue;
            }
        }
        __syncwarp();                                                          // lane 0's pushes are visible
        for (;;) {                                                             // backtrack
            if (level == 0) return 0;
            level--;
            const int fr = stack[level];
            cell = fr & 0xFF;
            const int qq = fr >> 8;
            if (qq == 0xFF) continue;                                          // advance was the last child
            const int s = qq < ns ? qq : qq - ns;
            const uint64_t ms = vw_shfl64(shape_at(sh, cell), s);
            if (qq < ns) { cnt4_sub(c, ms); tot_y++; if (lane == s) rem_y++; }
            else { cnt4_add(c, ms); tot_p++; if (lane == s) rem_p++; tot_mono += (int)((mono_mask >> s) & 1u); }
            q = qq + 1;
            break;
        }
        __syncwarp();                                                          // reads done before the slot is reused
    }
    return -1;
}

#if defined(VAL_PROFILE)
#define VW_CLOCK() clock64()
#else
#define VW_CLOCK() 0ll
#endif

// qw: lane l < 12 holds word l of the queue entry (env, puzzle, x | y << 8, pad, visited[8] dense), the others 0.
// Returns the SPARC_RULE_* flags (without SPARC_RULE_VALID), the same on every lane.
// stack: 128 uint16 of shared memory private to this warp (the exact-fit search's DFS stack).
// HARD = false (k_post): exact-fit searches run converged under `budget` nodes (0 = no limit); -1 is returned as soon as one outgrows it
//   (the snapshot is then validated again share by share, see k_val).
// HARD = true (one of VW_HARD_CODES shares of a hard item, this one running share `code` of every search): the poly
//   rule's verdict is left to the caller -- *fit_mask gets bit r when the r-th searched region has a fit in this share
//   (or another copy already reported one), *n_search the number of searched regions; the returned flags carry
//   SPARC_RULE_POLY_YLOP unless a region fails the area balance (and SPARC_RULE_ALL is left to the caller as well).
template <bool HARD>
__device__ __forceinline__ int validate_warp(const StoreDev& st, uint32_t qw, int lane, uint16_t* stack, int budget = 0, int code = 0,
                                             const unsigned long long* found = nullptr, unsigned long long* fit_mask = nullptr,
                                             int* n_search = nullptr) {
    unsigned long long fmask = 0;
    int nsearch = 0;
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
        uint32_t unl = C;                                                      // unlabelled rule cells, the same in all four groups
        const int w = (int)((h0 & 0xffu) - 1u) >> 1, hh = (int)(((h0 >> 8) & 0xffu) - 1u) >> 1;
        const int grp = lane >> 3;
        // Up to FOUR regions per round: the four 8-lane groups hold the same boards, so each floods from its own seed --
        // the unlabelled cells of rank 0, U/4, U/2, 3U/4 in raster order (spread out: neighbours tend to share a region).
        // Two groups that land in the same region just compute the same thing; rank 0 guarantees progress.
        for (;;) {
            const int pc = __popc(unl);                                        // <= 7 cells per lattice column
            const uint32_t packed = __reduce_or_sync(VW_FULL, (uint32_t)pc << (4 * k));
            const int U = (int)((((packed & 0x0F0F0F0Fu) + ((packed >> 4) & 0x0F0F0F0Fu)) * 0x01010101u) >> 24);
            if (!U) break;
            const uint32_t lowp = packed & ((1u << (4 * k)) - 1u);             // columns before mine
            const int below = (int)((((lowp & 0x0F0F0F0Fu) + ((lowp >> 4) & 0x0F0F0F0Fu)) * 0x01010101u) >> 24);
            const int j = ((grp * U) >> 2) - below;                            // my group's seed is cell j of my column, if 0 <= j < pc
            uint32_t f = 0;
            if (j >= 0 && j < pc) {
                uint32_t wv = unl;
#pragma unroll
                for (int i = 0; i < 6; i++) wv &= (i < j) ? wv - 1u : 0xFFFFFFFFu;
                f = wv & (0u - wv);
            }
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
                    ch = nw != f;                                              // (no change in the second sweep anywhere = fix-point)
                    f = nw;
                }
                if (!__any_sync(VW_FULL, ch)) break;
            }
            vp_flood += VW_CLOCK() - vp_tmp;
            // every lane gets all four components: word k of group j's flood, then its rule-cell board (bit 8cx + cy;
            // lane k holds column cx = k in the high half of its word)
            const uint32_t mine = f & C;
            uint64_t R[4];
            uint32_t all_cells = 0;
#pragma unroll
            for (int jg = 0; jg < 4; jg++) {
                const uint32_t cells = __shfl_sync(VW_FULL, mine, 8 * jg + k);
                all_cells |= cells;
                uint32_t t = (cells >> 17) & 0x5555u;
                t = (t | (t >> 1)) & 0x3333u; t = (t | (t >> 2)) & 0x0F0Fu; t = (t | (t >> 4)) & 0x00FFu;
                const uint32_t lo = __reduce_or_sync(VW_FULL, k < 4 ? t << (8 * k) : 0u);
                const uint32_t hi = __reduce_or_sync(VW_FULL, (k >= 4 && k < 7) ? t << (8 * (k - 4)) : 0u);
                R[jg] = vw_u64(lo, hi);
            }
            unl &= ~all_cells;
            // squares (SG.py:534-552): at most one distinct non-zero colour per region; stars (SG.py:554-614): every
            // star colour totals exactly 2 symbols in the region, no colourless star (SG.py:580-582).  Group = region,
            // lane = colour k; a region seen by two groups gets the same verdict twice.
            const uint64_t Rg = grp == 0 ? R[0] : grp == 1 ? R[1] : grp == 2 ? R[2] : R[3];
            const uint64_t rc = Rg & COLk;
            const unsigned vsq = __ballot_sync(VW_FULL, (SQ & rc) != 0ull);
            const unsigned vst = __ballot_sync(VW_FULL, ((STAR & rc) != 0ull && (__popcll(SB0 & rc) + 2 * __popcll(SB1 & rc) != 2)) ||
                                                        (STAR & Rg & ~ANYCOL) != 0ull);
            if (vst) star_ok = false;
#pragma unroll
            for (int jg = 0; jg < 4; jg++) if (__popc((vsq >> (8 * jg)) & 0xffu) > 1) sq_ok = false;
            // polys / ylops (SG.py:648-710), region after region with the whole warp: lane = shape slot
            if (n_shapes && poly_ok) {
                const uint64_t inst_cells = sh_poly | sh_ylop;
                const unsigned has = __reduce_or_sync(VW_FULL, ((inst_cells & R[0]) ? 1u : 0u) | ((inst_cells & R[1]) ? 2u : 0u) |
                                                               ((inst_cells & R[2]) ? 4u : 0u) | ((inst_cells & R[3]) ? 8u : 0u));
#pragma unroll
                for (int jg = 0; jg < 4; jg++) {
                    if (!((has >> jg) & 1u) || !poly_ok) continue;
                    bool dup = false;
#pragma unroll
                    for (int i = 0; i < jg; i++) dup = dup || R[i] == R[jg];
                    if (dup) continue;
                    const uint64_t Rj = R[jg];
                    const int pa = __popcll(sh_poly & Rj), yb = __popcll(sh_ylop & Rj);
                    const int net = __reduce_add_sync(VW_FULL, (pa - yb) * (int)(sh_anchors >> 56));
                    if (net != __popcll(Rj)) { poly_ok = false; continue; }    // SG.py:679-682
                    vp_tmp = VW_CLOCK();
                    if (hflags & 4) {              // more than six ylop instances: the plain search on 6-bit counters, every lane alike
                        Cnt8 np = {{0, 0, 0, 0}}, nyl = {{0, 0, 0, 0}};
                        for (unsigned mm = __ballot_sync(VW_FULL, (pa | yb) != 0); mm; mm &= mm - 1) {
                            const int sl = __ffs((int)mm) - 1;
                            cnt8_add(np, sl, __shfl_sync(VW_FULL, pa, sl)); cnt8_add(nyl, sl, __shfl_sync(VW_FULL, yb, sl));
                        }
                        const int fit = exact_fit_budget(Rj, w, hh, shapes, n_shapes, np, nyl, 0, true);
                        if (HARD) { const unsigned long long bit = 1ull << (nsearch & 63); nsearch++; if (fit) fmask |= bit; }
                        else if (!fit) poly_ok = false;
                    } else if (!HARD) {
                        const int fit = exact_fit_coop(Rj, w, hh, shapes, n_shapes, pa, yb, lane, stack, budget);
                        if (fit < 0) return -1;                                // a hard board: the snapshot becomes a hard item
                        if (!fit) poly_ok = false;
                    } else {
                        const unsigned long long bit = 1ull << (nsearch & 63);
                        nsearch++;
                        if (*reinterpret_cast<const volatile unsigned long long*>(found) & bit) fmask |= bit;
                        else {
                            Cnt8 np = {{0, 0, 0, 0}}, nyl = {{0, 0, 0, 0}};
                            for (unsigned mm = __ballot_sync(VW_FULL, (pa | yb) != 0); mm; mm &= mm - 1) {
                                const int sl = __ffs((int)mm) - 1;
                                cnt8_add(np, sl, __shfl_sync(VW_FULL, pa, sl)); cnt8_add(nyl, sl, __shfl_sync(VW_FULL, yb, sl));
                            }
                            if (exact_fit_code(Rj, w, hh, shapes, n_shapes, np, nyl, code, found, bit)) fmask |= bit;
                        }
                    }
                    vp_fit += VW_CLOCK() - vp_tmp;
                }
            }
        }
    }
    if (sq_ok) flags |= SPARC_RULE_SQUARES;
    if (star_ok) flags |= SPARC_RULE_STARS;
    if (poly_ok) flags |= SPARC_RULE_POLY_YLOP;
    if (HARD) { *fit_mask = fmask; *n_search = nsearch; }
    else if ((flags & 0xFF) == 0xFF) flags |= SPARC_RULE_ALL;                  // SG.py:932-935
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
