/* othello_b200 -- C ABI of the B200-native Othello self-play engine.
 *
 * This is the drop-in boundary under the reference's `othello_mcts` Python module
 * (reference: cpp/src/lib/othello_mcts.cpp:49-151).  Every entry point names the
 * reference interface it replaces.  Plain C types only, caller-owned buffers, no
 * exceptions across the boundary: every function returns an oz_status, and
 * oz_last_error() gives the message the reference would have thrown (same text).
 *
 * Unless a name ends in `_device`, pointer arguments are HOST buffers and the call does
 * the host<->device copies itself.  All compute runs in sm_100a kernels; there is no CPU
 * fallback -- without a usable CUDA device the compute calls return OZ_ERR_CUDA.
 * The exception is the scalar `oz_position_*` / `oz_get_*` value-type helpers, which are
 * the host-side Position of the boundary (reference: position.h:39-149 is host code
 * called one position at a time from Python) and share one __host__ __device__ source
 * (csrc/oz_rules.cuh) with the kernels.
 */
#ifndef OTHELLO_B200_H
#define OTHELLO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    OZ_OK = 0,
    OZ_ERR_INVALID_ARGUMENT = 1, /* reference: std::invalid_argument -> Python ValueError */
    OZ_ERR_OUT_OF_RANGE = 2,     /* reference: std::out_of_range     -> Python IndexError */
    OZ_ERR_CUDA = 3,             /* no device / launch failure (RuntimeError)              */
    OZ_ERR_CAPACITY = 4,         /* node arena or table exhausted (RuntimeError)           */
    OZ_ERR_STATE = 5             /* call order (RuntimeError)                              */
} oz_status;

/* Message for the last non-OK status on this thread. */
const char *oz_last_error(void);
/* ABI version; bumped on any signature change. */
int oz_abi_version(void);

/* ------------------------------------------------------------------ rules (position.h) */

/* reference: othello::Position, position.h:144-148 */
typedef struct {
    int32_t  player;     /* 1 Black, 2 White, 0 terminal */
    uint64_t p1;         /* Black discs; bit 63 = a1, square i <-> bit (63 - i) */
    uint64_t p2;         /* White discs */
    uint64_t legal;      /* legal moves of the side to move */
    uint64_t next_legal; /* legal moves of the other side, filled when `legal` is empty */
} oz_position;

/* Scalar value-type helpers (host).  reference: position.h:202-262, :264-272, :274-326,
 * checked variants :285-306 and :365-427 (same error texts). */
uint64_t oz_get_legal_moves(uint64_t player_discs, uint64_t opponent_discs);
uint64_t oz_get_flips(uint64_t move_mask, uint64_t player_discs, uint64_t opponent_discs);
int oz_position_initial(oz_position *out);
int oz_position_at(const oz_position *p, int index, int *square);             /* Position::at */
int oz_position_is_legal_move(const oz_position *p, int index, int *legal);   /* is_legal_move_checked */
int oz_position_legal_actions(const oz_position *p, int *actions /* >= 33 */, int *count);
int oz_position_apply_move(const oz_position *p, uint64_t move_mask, oz_position *out);  /* apply_move_checked */
int oz_position_apply_pass(const oz_position *p, oz_position *out);                     /* apply_pass_checked */
int oz_position_apply_action(const oz_position *p, int action, oz_position *out);       /* apply_action_checked */
/* UTF-8 board picture, position.h:429-456.  buf >= 512 bytes. */
int oz_position_to_string(const oz_position *p, char *buf, size_t buf_size);
int oz_transform_action(int action, int transformation);                       /* transformation.h:78-81 */

/* Batched rules on the GPU (host buffers; copies inside). */
int oz_movegen_batch(const uint64_t *player, const uint64_t *opponent, uint64_t *legal, size_t n);
int oz_flips_batch(const uint64_t *move, const uint64_t *player, const uint64_t *opponent,
                   uint64_t *flips, size_t n);
/* out[i] = in[i].apply_action(actions[i]) (unchecked, position.h:402-408). */
int oz_apply_action_batch(const oz_position *in, const int32_t *actions, oz_position *out, size_t n);
/* Same kernels on DEVICE pointers, asynchronous on `stream` (a cudaStream_t). */
int oz_movegen_batch_device(const uint64_t *player, const uint64_t *opponent, uint64_t *legal,
                            size_t n, void *stream);

/* Leaf count of the game tree `depth` plies below `root` (pass = one ply, terminal =
 * leaf).  mode 0 generates every leaf position; mode 1 counts the
 * last ply with popcount.  device_ms (optional) receives the CUDA-event time of the GPU
 * part.  The frontier is expanded breadth-first on the device, then one thread walks each
 * frontier position depth-first. */
int oz_perft(const oz_position *root, int depth, int mode, uint64_t *nodes, double *device_ms);

/* Measurement hook: issue rate of the integer pipe the move generator runs on (LOP3 + SHF, eight independent chains
 * per thread, nothing else in the loop), in thread-operations per second -- the roofline denominator of perft. */
int oz_debug_int_peak(double *ops_per_second, double *device_ms);

/* ------------------------------------------------------- encoding (transformation.h) */

/* features[i] = positions_to_features(history of item i, history_size, transformation[i])
 * (transformation.h:83-116).  history: n * history_size positions, item-major, index 0 =
 * current position; n_hist[i] <= history_size entries are valid.  features: n*(1+2H)*64
 * floats.  GPU kernel; host buffers. */
int oz_features_batch(const oz_position *history, const int32_t *n_hist, const int32_t *transformation,
                      int history_size, float *features, size_t n);

/* --------------------------------------------------------------- engine (mcts.h/.cpp) */

typedef struct oz_engine oz_engine;

enum { OZ_EVAL_EXTERNAL = 0, OZ_EVAL_SYNTH = 1, OZ_EVAL_HASH = 2, OZ_EVAL_RESNET = 3 };
enum { OZ_SYM_RANDOM = 0, OZ_SYM_IDENTITY = 1 };

/* reference: MCTS constructor kwargs, othello_mcts.cpp:90-112 / mcts.cpp:15-38.
 * torch_device / torch_pin_memory have no meaning below the ABI (the engine is always on
 * `device`); the Python layer keeps them for signature compatibility. */
typedef struct {
    int32_t  num_games;          /* lockstep games on this GPU (reference: always 1) */
    int32_t  history_size;       /* >= 1 */
    int32_t  num_simulations;    /* >= 1 */
    int32_t  num_threads;        /* >= 1; in-flight leaves per game = num_threads * batch_size */
    int32_t  batch_size;         /* >= 1 */
    float    c_puct_base;        /* > 0 */
    float    c_puct_init;        /* >= 0 */
    float    dirichlet_epsilon;  /* in [0, 1] */
    float    dirichlet_alpha;    /* >= 0 */
    uint64_t seed;               /* Philox key */
    uint32_t first_game_id;      /* global id of local game 0 (RNG streams are keyed by game id) */
    int32_t  device;             /* CUDA device ordinal */
    int32_t  evaluator;          /* OZ_EVAL_* */
    int32_t  sym_mode;           /* OZ_SYM_* (reference: uniform_int_distribution(0,7), search_thread.cpp:24) */
    int32_t  temperature_plies;  /* plies sampled proportionally to visits (train.py:422: 12) */
    int32_t  max_nodes_per_game; /* 0 = derive from num_simulations */
    float    self_play_temperature; /* > 0; sampling weights are visit_counts ** (1 / T) while ply < temperature_plies
                                       (train.py:423, --self-play-temperature; default 1) */
    int32_t  reserved;           /* 0 */
} oz_config;

void oz_config_default(oz_config *cfg);
int oz_engine_create(const oz_config *cfg, oz_engine **out);   /* MCTS::MCTS + validators mcts.cpp:167-241 */
void oz_engine_destroy(oz_engine *e);
/* Setter pairs of othello_mcts.cpp:131-150.  Same validation and messages. */
int oz_engine_set_history_size(oz_engine *e, int v);
int oz_engine_set_num_simulations(oz_engine *e, int v);
int oz_engine_set_num_threads(oz_engine *e, int v);
int oz_engine_set_batch_size(oz_engine *e, int v);
int oz_engine_set_c_puct_base(oz_engine *e, float v);
int oz_engine_set_c_puct_init(oz_engine *e, float v);
int oz_engine_set_dirichlet_epsilon(oz_engine *e, float v);
int oz_engine_set_dirichlet_alpha(oz_engine *e, float v);
/* OZ_EVAL_* used by oz_engine_search / oz_engine_play_ply. */
int oz_engine_set_evaluator(oz_engine *e, int evaluator);
int oz_engine_get_config(const oz_engine *e, oz_config *out);

/* Test hook: Dirichlet draws come from table[philox_word % len] instead of the on-device
 * gamma sampler (the reference's mt19937 stream cannot be seeded). len = 0 removes it. */
int oz_engine_set_gamma_table(oz_engine *e, const float *table, int len);

int oz_engine_reset(oz_engine *e);                                        /* MCTS::reset_position, all games */
int oz_engine_positions(oz_engine *e, oz_position *out /* num_games */);  /* MCTS::position */

/* Folded ResNet weights (neural_net.py:138-172) as one host blob; layout in
 * csrc/oz_resnet.cuh (oz_resnet_blob_bytes).  Copied to the device. */
typedef struct {
    int32_t in_planes;      /* 1 + 2 * history_size */
    int32_t channels;       /* conv_channels (multiple of 64) */
    int32_t blocks;         /* residual blocks */
    int32_t value_hidden;   /* value_head_hidden_channels */
} oz_resnet_shape;
size_t oz_resnet_blob_bytes(const oz_resnet_shape *shape);
/* Byte offsets of the blob sections, in the order documented in csrc/oz_resnet.cuh:
 * conv_w[0], conv_b[0], ..., conv_w[2*blocks], conv_b[2*blocks], policy_conv_w, policy_conv_b,
 * policy_fc_w, policy_fc_b, value_conv_w, value_conv_b, value_fc1_w, value_fc1_b, value_fc2_w,
 * value_fc2_b.  *count receives the number of offsets. */
int oz_resnet_blob_offsets(const oz_resnet_shape *shape, int64_t *offsets, int capacity, int *count);
int oz_engine_load_resnet(oz_engine *e, const oz_resnet_shape *shape, const void *blob, size_t bytes);

/* The network alone (reference: AlphaZeroNet.forward, neural_net.py:161-172, eval mode):
 * features [n, in_planes, 8, 8] float32 -> policy [n,65] softmax, value [n] tanh. */
typedef struct oz_resnet oz_resnet;
int oz_resnet_create(int device, oz_resnet **out);
void oz_resnet_destroy(oz_resnet *net);
int oz_resnet_load(oz_resnet *net, const oz_resnet_shape *shape, const void *blob, size_t bytes);
int oz_resnet_forward(oz_resnet *net, const float *features, size_t n, float *policy, float *value);
int oz_resnet_forward_device(oz_resnet *net, const float *features, size_t n, float *policy, float *value, void *stream);
/* Test hook: one 3x3 "same" convolution layer.  Host buffers: in NHWC bf16 [boards][8][8][cin],
 * w bf16 [9][cout][cin] (tap = ky*3+kx), bias f32 [cout], skip NHWC bf16 or NULL, out NHWC bf16.
 * cin % 32 == 0, cout % 128 == 0; boards is padded to 32 inside. */
int oz_debug_conv3x3(const uint16_t *in, int cin, const uint16_t *w, const float *bias, const uint16_t *skip,
                     int relu, uint16_t *out, int cout, size_t boards, double *device_ms);

/* Test / measurement hook: the native evaluator runs launch-bound batches of a 128-channel network (<= 32 boards per
 * 16-CTA cluster: the single-game MCTS.search of train.py:415-416 / player.py:251) as ONE cluster launch with the
 * activations in distributed shared memory (tower_cluster_kernel).  mode 0 forces the per-layer launches, 1 the tower
 * where it is eligible, -1 returns to the default (on; OZ_TOWER=0 in the environment turns it off).  Both paths issue the
 * same MMAs in the same order: results are bit-identical.  Applies to forwards launched (or graphs captured) afterwards. */
int oz_debug_set_tower(int mode);
/* Test / measurement hook: heads_kernel (neural_net.py:68-128) works on one board per warp while the batch is a single
 * wave of CTAs (<= 2 368 boards) and on four boards per warp beyond.  boards = 1 or 4 forces one of the two, anything else
 * returns to the choice by batch size (OZ_HEADS_BOARDS in the environment does the same).  Every output is the same sum in
 * the same order in both: results are bit-identical.  Applies to forwards launched (or graphs captured) afterwards. */
int oz_debug_set_heads_boards(int boards);
/* Measurement hook for the same kernel.  *max_clusters (optional) = how many 16-CTA clusters the current device runs at once
 * (0: the tower cannot be used here).  out (optional): [64][8] int64 clock64 stamps of cluster rank 8 during the last tower
 * launch, per layer: first MMA issued, all MMAs issued, accumulator seen by the epilogue, neighbours' MMAs seen, first / second
 * 32-channel half written, then (next layer's hand-over) own rows seen, halo rows seen.  Recorded only when OZ_TOWER_DBG is set
 * in the environment; profiles/r02_tower_timeline.txt is such a record. */
int oz_debug_tower_timeline(long long *out, int *max_clusters);

/* Test hook: weight gradient of one 3x3 "same" convolution layer (the training step's tcgen05 kernel, csrc/oz_wgrad.cu).
 * Host buffers: x NHWC bf16 [boards][8][8][cinp] (cinp = 32 or a multiple of 128), dz NHWC bf16 [boards][8][8][cout]
 * (cout a multiple of 128), boards a multiple of 16; G fp32 [cout][9][cinp] (tap = ky*3+kx) = sum over positions of
 * dz[p][co] * x[p + tap][ci]. */
int oz_debug_wgrad3x3(const uint16_t *x, int cinp, const uint16_t *dz, int cout, size_t boards, float *G, double *device_ms);

/* MCTS::search (mcts.h:220-256) for every game, with a built-in evaluator
 * (synth / hash / resnet).  Blocking. */
int oz_engine_search(oz_engine *e);
/* The same search for the games [first_game, first_game + num_games) only; the others keep their trees untouched and
 * cost nothing (the network batch is num_games * T * B).  New capability for lockstep match play, where each of two
 * engines (one per network) searches the games in which its player is to move (reference: evaluation.py:15-90 plays
 * such games one at a time through play_game, player.py:33-96). */
int oz_engine_search_range(oz_engine *e, int first_game, int num_games);
/* The same search split at the evaluator, for OZ_EVAL_EXTERNAL (reference: the Python
 * callable of othello_mcts.cpp:36-45).  rounds() = ceil(N / (T*B)).  Each round:
 * select_device writes features [num_games*T*B, 1+2H, 8, 8] float32 to the DEVICE buffer,
 * the caller evaluates them, expand_device consumes policy [.,65] and value [.] DEVICE
 * buffers.  All on `stream`. */
int oz_engine_search_begin(oz_engine *e);
int oz_engine_rounds(const oz_engine *e);
int oz_engine_select_device(oz_engine *e, float *features, void *stream);
int oz_engine_expand_device(oz_engine *e, const float *policy, const float *value, void *stream);
int oz_engine_search_end(oz_engine *e);

/* MCTS::visit_counts / mean_action_values (mcts.cpp:45-61) for every game: row g holds
 * n_children[g] entries in legal_actions() order.  stride = 36.  Any pointer may be NULL. */
#define OZ_MAX_CHILDREN 36
int oz_engine_root_stats(oz_engine *e, int32_t *n_children, int32_t *visit_counts,
                         float *mean_action_values, int32_t *root_visit_count);
/* MCTS::self_play_data (mcts.cpp:63-112) of one game: features [8,1+2H,8,8], policy [8,65]. */
int oz_engine_self_play_data(oz_engine *e, int game, float *features, float *policy);
/* MCTS::apply_action (mcts.cpp:114-165) for every game; actions[g] < 0 leaves game g alone.
 * Re-roots, drops the siblings, keeps the ancestors for the history planes. */
int oz_engine_apply_actions(oz_engine *e, const int32_t *actions);
/* train.py:421-430 on the device with the engine's counter-based stream: proportional to
 * visit counts while ply < temperature_plies, then a random argmax.  -1 for finished games. */
int oz_engine_choose_actions(oz_engine *e, int32_t *actions);

/* One self-play ply for every unfinished game: search, choose, record, apply (train.py:415-436).
 * Outputs (optional, host): chosen actions and root visit counts [num_games * OZ_MAX_CHILDREN].
 * *unfinished receives the number of games still running afterwards. */
int oz_engine_play_ply(oz_engine *e, int32_t *actions, int32_t *visit_counts, int32_t *unfinished);

/* Compact game records written by play_ply: one row per (game, ply). */
typedef struct {
    uint64_t p1, p2;
    int32_t  player;
    int32_t  action;
    int32_t  n_children;
    int32_t  visit_counts[OZ_MAX_CHILDREN];
} oz_ply_record;
/* plies[g] = plies played by game g; records: num_games * max_plies rows (game-major). */
int oz_engine_records(oz_engine *e, int32_t *plies, oz_ply_record *records, int max_plies);
/* Expand records into the training tensors _self_play returns (train.py:432-450): for every
 * recorded ply of every finished game, 8 symmetries of features [1+2H,8,8], policy [65] and
 * the outcome value.  Returns the sample count in *n_samples; buffers may be NULL to query. */
int oz_engine_training_samples(oz_engine *e, float *features, float *policy, float *value,
                               size_t capacity, size_t *n_samples);

/* The same expansion for compact records that travelled (a multi-GPU run gathers 176-byte oz_ply_record rows and expands
 * them on the destination rank instead of shipping 4 612-byte samples): records [n_games][max_plies] + plies [n_games]
 * as oz_engine_records / oz_engine_finished_records return them; every game counts as finished (its outcome follows from
 * the last record and action).  Host buffers; the kernel runs on `device`.  Buffers may be NULL to query *n_samples. */
int oz_training_samples_from_records(int device, const oz_ply_record *records, const int32_t *plies, int n_games, int max_plies,
                                     int history_size, float *features, float *policy, float *value, size_t capacity,
                                     size_t *n_samples);

/* Rolling game pool (new capability; the reference replays one MCTS object game after game, train.py:394-410):
 * the engine's num_games slots play `total_games` games with ids first_game_id .. first_game_id + total_games - 1.
 * After every oz_engine_play_ply, finished games move to a store of completed games (completion order) and their
 * slots restart from the initial position with the next unplayed id, so the leaf batch stays full until the pool
 * runs dry.  RNG streams are keyed by game id: a game's record does not depend on the slot or GPU that played it.
 * total_games = 0 switches the pool off.  Resets all games.  With a pool, oz_engine_training_samples covers the
 * completed games of the store. */
int oz_engine_set_game_pool(oz_engine *e, int total_games);
int oz_engine_finished_games(oz_engine *e, int32_t *count);
/* Rows 0 .. min(*count, capacity) - 1 of the store: global game id, plies played, records [row][max_plies]. */
int oz_engine_finished_records(oz_engine *e, uint32_t *game_ids, int32_t *plies, oz_ply_record *records, int max_plies,
                               int capacity, int32_t *count);

typedef struct {
    uint64_t simulations;      /* reference accounting: rounds * T * B per search per game */
    uint64_t evaluations;      /* network slots actually evaluated */
    uint64_t nodes_expanded;
    uint64_t searches;
    uint64_t plies;
    uint64_t games_finished;
    uint64_t kernel_launches;  /* launches of this library's kernels */
    double   select_ms, eval_ms, expand_ms, reroot_ms; /* CUDA-event totals when profiling is on */
    double   conv_ms;          /* part of eval_ms spent in the 3x3 convolution launches */
    uint64_t conv_launches;    /* convolution launches timed in conv_ms */
    uint64_t path_nodes;       /* tree levels descended by the selections (sum of path lengths, root excluded) */
    uint64_t children_read;    /* child (N, Q, P, header) rows read by PUCT selection */
} oz_counters;
int oz_engine_counters(oz_engine *e, oz_counters *out);
int oz_engine_set_profiling(oz_engine *e, int on);
/* CUDA-graph replay of a search (all its rounds in one graph, kernels chained by programmatic dependent launch): 0 never,
 * 1 (default) when the leaf batch is launch-bound (<= 8192 boards, e.g. the single-game MCTS of the reference's play CLI),
 * 2 always.  Results are identical either way. */
int oz_engine_set_graphs(oz_engine *e, int mode);
/* Pipelined search (new capability; the reference evaluates one 16-leaf batch at a time on one thread, mcts.h:220-256):
 * the games are split into `groups` contiguous groups that run select -> network -> expand on their own streams, so one
 * group's tree kernels run underneath another group's convolutions.  groups: 0 = automatic, 1 = off, 2..8.  Automatic is
 * off for the eager searches of large batches (on B200 the convolutions lose more to the co-running tree kernels than the
 * hidden tree time gives back) and 8 groups for graph-replayed searches of 8 to 48 games of a 128-channel network, where the
 * groups are parallel BRANCHES of the search's graph, each with its own cluster-tower launches (8 games of 128x10b: 8.6 ->
 * 6.7 ms per ply).  pdl != 0 keeps programmatic dependent launch between the conv layers while pipelined eagerly.
 * Results are identical for every setting (every game's search is independent of the others). */
int oz_engine_set_pipeline(oz_engine *e, int groups, int pdl);
/* Stream the engine launches on (a cudaStream_t), for callers that time with CUDA events. */
void *oz_engine_stream(oz_engine *e);

/* ------------------------------------------------------------ training step (train.py:486-502) */

/* One optimisation step of the reference's _train loop on the GPU: forward in train mode (BatchNorm batch statistics,
 * running statistics updated), policy cross-entropy + value MSE + explicit L2 over all parameters (train.py:494-499),
 * backward, torch.optim.SGD with momentum (train.py:261).  The 3x3 convolutions (forward and input gradients) run on the
 * engine's tcgen05 kernel in bf16 with fp32 accumulation; BatchNorm, heads, losses and the optimizer are fp32; master
 * weights are fp32.  Parameters travel as ONE flat float vector in the order of AlphaZeroNet.parameters()
 * (= torch.nn.utils.parameters_to_vector); buffers as the running_mean, running_var of every BatchNorm in module order. */
typedef struct oz_trainer oz_trainer;
typedef struct {
    float lr;          /* SGD learning rate (train.py:103-106) */
    float momentum;    /* SGD momentum (train.py:109-112) */
    float l2_weight;   /* --l2-weight-regulation (train.py:202-206) */
    float bn_momentum; /* nn.BatchNorm2d momentum (0.1) */
    float bn_eps;      /* nn.BatchNorm2d eps (1e-5) */
} oz_train_config;
void oz_train_config_default(oz_train_config *cfg);
/* max_batch: largest batch size (a multiple of 32; batches must be multiples of 32 as well). */
int oz_trainer_create(int device, const oz_resnet_shape *shape, int max_batch, oz_trainer **out);
void oz_trainer_destroy(oz_trainer *t);
size_t oz_trainer_num_params(const oz_trainer *t);
size_t oz_trainer_num_buffers(const oz_trainer *t);
/* Host vectors; any pointer may be NULL (left as it is / not read back).  momentum = the SGD momentum buffers. */
int oz_trainer_set_state(oz_trainer *t, const float *params, const float *buffers, const float *momentum);
int oz_trainer_get_state(oz_trainer *t, float *params, float *buffers, float *momentum, float *grads);
/* features [n][in_planes][8][8], policy targets [n][65], value targets [n] (HOST float32, what the DataLoader of
 * train.py:472-489 yields).  apply_update = 0 computes losses and gradients only.  losses (optional, host):
 * total, policy, value, l2 -- the four numbers train.py:504-507 logs. */
int oz_trainer_step(oz_trainer *t, const float *features, const float *policy, const float *value, int n,
                    const oz_train_config *cfg, int apply_update, float *losses);
/* The same on DEVICE pointers (asynchronous on the trainer's stream unless losses != NULL). */
int oz_trainer_step_device(oz_trainer *t, const float *features, const float *policy, const float *value, int n,
                           const oz_train_config *cfg, int apply_update, float *losses);
/* Network outputs of the last step's forward pass (train mode): policy [n][65], value [n]. */
int oz_trainer_outputs(oz_trainer *t, int n, float *policy, float *value);
/* CUDA-graph replay of the whole step (default on): a (batch size, hyper-parameters) combination runs eagerly the first
 * time, is captured the second time and replayed afterwards.  Results are identical either way. */
int oz_trainer_set_graphs(oz_trainer *t, int on);
void *oz_trainer_stream(oz_trainer *t);

#ifdef __cplusplus
}
#endif
#endif
