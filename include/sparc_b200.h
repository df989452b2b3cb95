/*
 * sparc_b200.h -- C ABI of the B200-native batched SPaRC environment (libsparc_b200.so).
 *
 * This is the drop-in boundary for the reference's hot path
 *     reset -> step -> (terminal) rule validation -> one-hot 'new' observation
 * (reference: SPaRC_Gym/SPaRC_Gym.py, "SG.py" below).  The reference is pure Python and has
 * no FFI; each entry point names the Python method it replaces.  The host side that mirrors
 * the reference's Gymnasium surface (sparc_gym_b200/vec_env.py, gym_env.py) binds these with
 * ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (SPARC_E_*); the message is available
 *     from sparc_last_error() (thread-local).  Nothing throws across the boundary.
 *   - all *_dev pointers are DEVICE pointers owned by the caller (e.g. torch tensors'
 *     data_ptr()); the library never allocates or frees them.  The library owns the puzzle
 *     store and the per-env state.
 *   - all work is enqueued on the caller's CUDA stream (`stream` is a cudaStream_t passed as
 *     void*); there is no hidden host synchronisation except in the *_host convenience calls,
 *     which say so.
 *   - one handle <-> one device; a handle is not thread-safe; distinct handles are independent.
 *   - there is NO CPU fallback: without a CUDA device every call fails with SPARC_E_CUDA.
 */
#ifndef SPARC_B200_H
#define SPARC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPARC_ABI_VERSION 1

/* geometry limits: node lattice (2w+1) x (2h+1) with w,h <= 7 (SG.py:243-248) */
#define SPARC_MAX_DIM        15
#define SPARC_BOARD_WORDS     8      /* 256-bit node boards */
#define SPARC_MAX_PLANES     10      /* visited,gaps,agent_location,target_location + 6 symbol kinds */
#define SPARC_MAX_SHAPES     32      /* distinct polyshapes referenced by one puzzle */
#define SPARC_MAX_YLOPS       6      /* ylop instances in one puzzle served by the 4-bit demand counters of the fast search */
#define SPARC_MAX_YLOPS_WIDE 30      /* ... by the 6-bit counters (header flag bit 2); the reference has no limit, but its search
                                        enumerates 49^k ylop anchors (SG.py:794-816) */
#define SPARC_MAX_REGIONS    64      /* region ids per env (one per rule cell at most: 7 x 7 = 49) */
/* poly/ylop verdict of one region (SG.py:674-702) */
enum { SPARC_FIT_NONE = 0,           /* no poly / ylop instance in the region */
       SPARC_FIT_AREA_MISMATCH = 1,  /* sum(poly) - sum(ylop) != region area: exact fit skipped */
       SPARC_FIT_NO_TILING = 2,      /* area matches, no exact placement exists */
       SPARC_FIT_OK = 3 };

/* error codes */
#define SPARC_OK              0
#define SPARC_E_ARG          -1
#define SPARC_E_CUDA         -2
#define SPARC_E_NOMEM        -3

/* canonical plane ids of the device-side plane superset (the per-puzzle key ORDER of the
 * reference, SG.py:191-196 / 305-306, is restored on the host by the N=1 drop-in) */
enum { SPARC_PL_VISITED = 0, SPARC_PL_GAPS = 1, SPARC_PL_AGENT = 2, SPARC_PL_TARGET = 3,
       SPARC_PL_STAR = 4, SPARC_PL_SQUARE = 5, SPARC_PL_TRIANGLE = 6, SPARC_PL_POLY = 7,
       SPARC_PL_YLOP = 8, SPARC_PL_DOT = 9 };

/* reward codes (SG.py:1204-1223): value = {0, +0.01, -0.01, +1, -1} */
enum { SPARC_R_ZERO = 0, SPARC_R_ON_PATH = 1, SPARC_R_OFF_PATH = 2, SPARC_R_WIN = 3, SPARC_R_LOSE = 4 };

/* rule flag bits, in the order of SG.py:912-935 */
enum { SPARC_RULE_REACHED_TARGET = 1 << 0, SPARC_RULE_PATH_NOT_CROSSING = 1 << 1,
       SPARC_RULE_NO_GAP_VIOLATIONS = 1 << 2, SPARC_RULE_ALL_DOTS = 1 << 3,
       SPARC_RULE_SQUARES = 1 << 4, SPARC_RULE_STARS = 1 << 5, SPARC_RULE_TRIANGLES = 1 << 6,
       SPARC_RULE_POLY_YLOP = 1 << 7, SPARC_RULE_ALL = 1 << 8,
       SPARC_RULE_VALID = 1 << 15 /* set when the flags were computed for this env at this step */ };

/* autoreset modes (no reference counterpart: the reference is single-env and never guards a
 * finished episode) */
enum { SPARC_AUTORESET_DISABLED = 0, SPARC_AUTORESET_NEXT_STEP = 1, SPARC_AUTORESET_SAME_STEP = 2,
       SPARC_AUTORESET_CONTINUE = 3 /* the reference's own behaviour (SG.py:1131-1238 has no guard): a finished
                                       episode is stepped like a live one -- the agent keeps moving, every further step
                                       returns +-1.  Used by the single-env drop-in; episode statistics then count every
                                       such step as an episode end. */ };
/* when the rule validator runs inside sparc_step */
enum { SPARC_VALIDATE_NEVER = 0, SPARC_VALIDATE_TERMINAL = 1, SPARC_VALIDATE_EVERY_STEP = 2 };
/* action dtypes accepted by sparc_step */
enum { SPARC_ACT_U8 = 0, SPARC_ACT_I32 = 1, SPARC_ACT_I64 = 2 };

/* episode statistics, int64[8] (fixed point keeps the multi-GPU sum exact) */
enum { SPARC_STAT_EPISODES = 0, SPARC_STAT_SOLVED = 1, SPARC_STAT_FAILED = 2, SPARC_STAT_TRUNCATED = 3,
       SPARC_STAT_STEPS = 4, SPARC_STAT_RETURN_CENTS = 5, SPARC_STAT_LENGTH = 6, SPARC_STAT_RULES_PASS = 7,
       SPARC_N_STATS = 8 };

/* ---- puzzle store (replaces the list of dicts built by _process_puzzles, SG.py:219-368) ---- */

typedef struct {                 /* 32 bytes */
    uint8_t  X, Y;               /* node-lattice size of this puzzle (<= slab) */
    uint8_t  start_x, start_y;   /* SG.py:267 */
    uint8_t  target_x, target_y; /* SG.py:268 */
    uint8_t  n_shapes;           /* shape slots of this puzzle */
    uint8_t  flags;              /* bit0: solution_count > 0; bit1: has stars / squares / shape instances
                                    (a region rule can fail, so the validator must segment regions);
                                    bit2: more than SPARC_MAX_YLOPS ylop instances (wide demand counters) */
    uint32_t shape_off;          /* index of the first SparcShape */
    uint32_t trie_root;          /* trie node of the start location, 0 = no listed solution starts there */
    int32_t  solution_count;
    int32_t  difficulty;
    uint32_t reserved[2];
} SparcPuzzleHeader;

/* per-puzzle node boards (8 x uint32 each).  Board 0 is in dense slab order (bit = x*slab_y+y),
 * the others in 16-stride order (bit = x*16+y) used by the validator. */
enum { SPARC_NB_WALK = 0,      /* in-bounds and gaps==0 : where the agent may stand (SG.py:1039) */
       SPARC_NB_GAPS16 = 1,    /* gaps plane incl. all rule cells (SG.py:345-351) */
       SPARC_NB_CELL16 = 2,    /* rule cells (odd,odd) (SG.py:374-388) */
       SPARC_NB_VALID16 = 3,   /* x < X and y < Y */
       SPARC_NB_DOT16 = 4,     /* dot plane (SG.py:521-532) */
       SPARC_NB_TRI1 = 5, SPARC_NB_TRI2 = 6, SPARC_NB_TRI3 = 7, SPARC_NB_TRI4 = 8, /* interior triangles by count */
       SPARC_NB_TRIBAD = 9,    /* interior triangles whose count can never match (>4) (SG.py:627-635) */
       SPARC_N_NODE_BOARDS = 10 };
/* per-puzzle cell boards (uint64, bit = cx*8+cy for rule cell (2cx+1, 2cy+1)) */
enum { SPARC_CB_CELLS = 0, SPARC_CB_STAR = 1, SPARC_CB_SQUARE = 2,
       SPARC_CB_SYM1 = 3,      /* bit 0 of the number of symbol layers the cell is in (SG.py:468-482) */
       SPARC_CB_SYM2 = 4,      /* bit 1 of that number (a typed cell that is also a dot counts twice) */
       SPARC_CB_COLOR1 = 5,    /* ... COLOR1+7 : cells of colour 1..8 (SG.py:310) */
       SPARC_N_CELL_BOARDS = 16 };

typedef struct {                 /* 32 bytes: one 16-byte load fetches what a placement needs */
    uint64_t mask;               /* bits 0..55: shape cells, bit = dx*8 + (dy - min_dy); anchor = lexicographically
                                    first filled entry, array index 0 -> grid x (SG.py:856-871);
                                    bits 56..58: anchor_dy = -min_dy, the anchor's column inside the mask;
                                    bits 59..61: flags -- bit0: larger than any 7x7 board, bit1: no filled entry,
                                    bit2: the filled entries are 4-connected */
    uint64_t anchors;            /* bits 0..55: cells (bit = 8cx + cy) of THIS puzzle's board on which the anchor may
                                    sit with every shape cell in bounds (SG.py:880-883); bits 56..63: area = sum of
                                    the shape array (SG.py:733) */
    uint64_t poly_cells;         /* cells holding a 'poly' instance of this shape (SG.py:716-736) */
    uint64_t ylop_cells;         /* cells holding a 'ylop' instance */
} SparcShape;

/* trie of the first `solution_count` listed solution paths (SG.py:1205-1223, 1244-1265):
 * 8 x uint32 per node = child[4] (by action id), parent, is_end, depth, pad; node 0 = null */
#define SPARC_TRIE_NODE_WORDS 8

typedef struct {                 /* all pointers are HOST pointers; copied by sparc_store_create */
    int32_t n_puzzles;
    int32_t slab_x, slab_y;      /* common padded lattice of the pool (>= every X, Y) */
    int32_t n_planes;            /* planes in the pool's superset, visited first */
    int32_t plane_ids[SPARC_MAX_PLANES];   /* SPARC_PL_* of each plane, in output order */
    int32_t words_per_puzzle;    /* ceil(n_planes*slab_x*slab_y / 32) */
    int32_t n_shapes_total;
    int32_t n_trie_nodes;
    const SparcPuzzleHeader* headers;   /* [n_puzzles] */
    const uint32_t* node_boards;        /* [n_puzzles][SPARC_N_NODE_BOARDS][8] */
    const uint64_t* cell_boards;        /* [n_puzzles][SPARC_N_CELL_BOARDS] */
    const uint32_t* static_bits;        /* [n_puzzles][words_per_puzzle]: the observation's one-hot
                                           planes in output order, bit-packed, dynamic planes zero */
    const uint32_t* color4;             /* [n_puzzles][ceil(4*slab_x*slab_y/32)]: colour ids 0..8, 4 bits per
                                           node, dense slab order  (SG.py:279, 309-315) */
    const uint32_t* info_nz;            /* [n_puzzles][ceil(slab_x*slab_y/32)]: additional_info != 0 */
    const int32_t*  info;               /* [n_puzzles][slab_x*slab_y]  (SG.py:280, 318-325) */
    const SparcShape* shapes;           /* [n_shapes_total] */
    const uint32_t* trie;               /* [n_trie_nodes][8] */
} SparcStoreDesc;

typedef struct SparcStore SparcStore;
typedef struct SparcEnv SparcEnv;

/* observation 'new' (SG.py:956-979), written in full every call (update = SPARC_OBS_REWRITE).
 * format SPARC_OBS_REFERENCE (0, the reference's dtypes -- the bit-exact drop-in and the headline benchmark):
 *   base  int32 [n_envs][n_planes][slab_x][slab_y]   (one-hot planes, SG.py:272-277, 305-306)
 *   color int32 [n_envs][slab_x][slab_y]             (SG.py:279)
 *   info  int64 [n_envs][slab_x][slab_y]             (SG.py:280)
 * format SPARC_OBS_COMPACT (1, same values in the narrowest types that hold them -- 3.5x fewer bytes for a
 * policy that lives on the GPU; reported under its own algorithmic bytes, never mixed with the above):
 *   base  uint8 [..], color uint8 [..], info int32 [..]   (same shapes)
 * contiguous, 16-byte aligned; any pointer may be NULL to skip that array. */
enum { SPARC_OBS_REFERENCE = 0, SPARC_OBS_COMPACT = 1 };
/* update SPARC_OBS_DELTA (sparc_step only): the arrays are the env's PERSISTENT observation, as in the reference,
 * whose obs dict aliases the live planes and mutates them in place (SG.py:979, 1146-1188): a step touches only
 * what changed -- visited / agent_location at the cell left and the cell entered -- and rewrites an env in full
 * only when it starts a new episode (autoreset).  Same contents as a full rewrite after every call, provided the
 * caller passes the same arrays every step and they hold the env's current observation (sparc_reset and
 * sparc_encode always write in full).  A separate mode with its own algorithmic bytes, never mixed into the
 * full-rewrite figures. */
enum { SPARC_OBS_REWRITE = 0, SPARC_OBS_DELTA = 1 };
typedef struct {
    void*   base_dev;
    void*   color_dev;
    void*   info_dev;
    int32_t format;      /* SPARC_OBS_REFERENCE / SPARC_OBS_COMPACT */
    int32_t update;      /* SPARC_OBS_REWRITE / SPARC_OBS_DELTA */
} SparcObsOut;

/* per-env step results (SG.py:1238); any pointer may be NULL */
typedef struct {
    float*    reward_dev;        /* 0, +-0.01f, +-1 */
    int8_t*   reward_code_dev;   /* SPARC_R_* : exact */
    uint8_t*  terminated_dev;    /* SG.py:1192 */
    uint8_t*  truncated_dev;     /* SG.py:1134, 1195-1199 */
    uint8_t*  legal_mask_dev;    /* bit a = action a legal in the returned state (SG.py:1024-1051) */
    uint16_t* rule_flags_dev;    /* SPARC_RULE_* | SPARC_RULE_VALID when validated this step, else 0 */
    int32_t*  step_count_dev;    /* info['current_step'] (SG.py:1017) */
    uint8_t*  agent_xy_dev;      /* [n_envs][2] info['agent_location'] (SG.py:1018) */
} SparcStepOut;

int         sparc_abi_version(void);
const char* sparc_last_error(void);

/* replaces SPaRC_Gym.__init__ -> _process_puzzles (SG.py:77-90, 219-368): upload the pool */
int sparc_store_create(const SparcStoreDesc* desc, int device, SparcStore** out);
int sparc_store_destroy(SparcStore* store);

/* replaces the per-env state of SPaRC_Gym (SG.py:63-66, 141-187): n_envs lockstep instances.
 * puzzle_stride: on autoreset an env moves from puzzle p to (p + puzzle_stride) mod n_puzzles
 * (1 mirrors the reference's cyclic cursor, SG.py:1087; 0 replays the same puzzle). */
int sparc_env_create(SparcStore* store, int64_t n_envs, int traceback, int max_steps,
                     int autoreset_mode, int puzzle_stride, int validate_mode, SparcEnv** out);
int sparc_env_destroy(SparcEnv* env);

/* replaces reset() (SG.py:1057-1108) with a CLEAN reset.  puzzle_idx_dev: int32[n_envs]
 * puzzle index per env (taken modulo n_puzzles, Python style: -1 is the last puzzle; the indices live on
 * the device, so a range error cannot be reported without a sync -- host wrappers check before uploading),
 * or NULL = keep the env's current puzzle.  A freshly created env is already a clean reset on puzzle
 * i mod n_puzzles.  Every entry point works on the store's device and leaves the calling thread's current device as it was.
 * mask_dev: uint8[n_envs] which envs to reset, or NULL = all.  Writes the observation of all envs. */
int sparc_reset(SparcEnv* env, const int32_t* puzzle_idx_dev, const uint8_t* mask_dev,
                const SparcStepOut* info_out, const SparcObsOut* obs, void* stream);

/* replaces step() (SG.py:1111-1238) for all envs in lockstep. */
int sparc_step(SparcEnv* env, const void* actions_dev, int action_dtype,
               const SparcStepOut* out, const SparcObsOut* obs, void* stream);

/* replaces _get_obs() (SG.py:956-979) on its own: rewrite the observation of every env from the
 * current state without stepping (after sparc_state_load, or to fill a second buffer). */
int sparc_encode(SparcEnv* env, const SparcObsOut* obs, void* stream);

/* replaces _validate_rules() (SG.py:941-950) on the current state of the masked envs (NULL = all).
 * flags_dev: uint16[n_envs]; region_map_dev: int8[n_envs][slab_x][slab_y] (-1 off-cell,
 * SG.py:429-454) or NULL; region_fit_dev: int8[n_envs][SPARC_MAX_REGIONS] or NULL -- the poly/ylop
 * verdict SPARC_FIT_* of every region id (the `region_details` of SG.py:674-702). */
int sparc_validate(SparcEnv* env, const uint8_t* mask_dev, uint16_t* flags_dev,
                   int8_t* region_map_dev, int8_t* region_fit_dev, void* stream);

/* episode statistics accumulated since creation / the last clear: copies int64[SPARC_N_STATS]
 * to out_dev (device) on `stream`; clear != 0 zeroes the accumulators afterwards. */
int sparc_stats(SparcEnv* env, int64_t* out_dev, int clear, void* stream);

/* HOST-buffer step, the end-to-end call: takes `actions_host` (uint8[n_envs], pageable or pinned), steps, and
 * fills the host arrays reward / terminated / truncated (any may be NULL).  Observations stay on the device in
 * `obs`.  The call returns when the RESULTS are in the host arrays: the observation rewrite of this step may still
 * be running then; it is ordered before anything put on `stream` later (kernels, copies, a synchronise), so a
 * device-side consumer needs no extra care and a host-side one synchronises `stream` first.
 * Launch-bound batches (<= 8,192 envs) are one launch with the actions in the kernel parameters and a completion
 * flag in mapped host memory; bigger ones upload / download on a side stream beside the encoder.
 * SPARC_HOST_EARLY=0 / SPARC_HOST_ZERO_COPY=0 in the environment restore the plain copy + synchronise sequences. */
int sparc_step_host(SparcEnv* env, const uint8_t* actions_host, float* reward_host,
                    uint8_t* terminated_host, uint8_t* truncated_host,
                    const SparcObsOut* obs, void* stream);

/* HOST-record step / reset: what the single-env Gymnasium drop-in needs after every call, in ONE device->host copy
 * (the reference's step() returns info with legal actions, step counter, agent location and the rule status of the
 * CURRENT state, SG.py:994-1022): steps (or resets) all envs, validates every env's current state (flags, region
 * map, per-region poly/ylop verdicts: sparc_validate) and fills `record_host` with the record below; like
 * sparc_step_host it returns when the record is there (up to 16 envs: ONE launch, arguments in the kernel parameters,
 * completion flag in mapped host memory; more: launches + copy + synchronise).  actions_host / puzzle_idx_host: int32[n_envs] in host memory (puzzle_idx_host may be NULL = keep).
 * Record, struct of arrays over the n envs, sparc_info_record_bytes(env) bytes:
 *   int32 step[n] | uint16 flags[n] (SPARC_RULE_* | SPARC_RULE_VALID) | uint8 agent_xy[n][2] | int8 reward_code[n] |
 *   uint8 terminated[n] | uint8 truncated[n] | uint8 legal_mask[n] | int8 region_fit[n][SPARC_MAX_REGIONS] |
 *   int8 region_map[n][slab_x][slab_y] */
int64_t sparc_info_record_bytes(const SparcEnv* env);
int sparc_step_info_host(SparcEnv* env, const int32_t* actions_host, void* record_host, const SparcObsOut* obs, void* stream);
int sparc_reset_info_host(SparcEnv* env, const int32_t* puzzle_idx_host, void* record_host, const SparcObsOut* obs, void* stream);

/* checkpoint / restore of the per-env state: opaque blob of
 * sparc_state_bytes(env) bytes, device memory. */
int64_t sparc_state_bytes(const SparcEnv* env);
int sparc_state_save(SparcEnv* env, void* blob_dev, void* stream);
int sparc_state_load(SparcEnv* env, const void* blob_dev, void* stream);

/* introspection */
int64_t sparc_env_num_envs(const SparcEnv* env);
int     sparc_kernel_launches(const SparcEnv* env, int64_t* out_count);  /* launches issued so far */

#ifdef __cplusplus
}
#endif
#endif /* SPARC_B200_H */
