// Native ResNet evaluator (reference: python/othello_alphazero/neural_net.py:138-172):
// folded-BN bf16 3x3 convolutions as implicit GEMM on tcgen05 / TMEM fed by TMA, fused
// bias + residual + ReLU epilogue, and one fused kernel for both heads.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "oz_common.cuh"

namespace oz {

// Host blob layout (all offsets 256-byte aligned), produced by the Python weight folder
// (othello_mcts/weights.py) and mirrored by oz_resnet_blob_offsets():
//   conv[0]            bf16 [9][C][32]   tap-major, cout, cin (planes zero-padded to 32)
//   bias[0]            f32  [C]
//   conv[1..2*blocks]  bf16 [9][C][C]
//   bias[1..2*blocks]  f32  [C]
//   policy_conv  f32 [2][C]   policy_conv_bias f32 [2]
//   policy_fc    f32 [65][128] policy_fc_bias  f32 [65]
//   value_conv   f32 [C]      value_conv_bias  f32 [1]
//   value_fc1    f32 [Hv][64] value_fc1_bias   f32 [Hv]
//   value_fc2    f32 [Hv]     value_fc2_bias   f32 [1]
struct ResNetLayout {
    std::vector<size_t> conv_w, conv_b;  // per conv layer
    size_t pconv_w, pconv_b, pfc_w, pfc_b, vconv_w, vconv_b, vfc1_w, vfc1_b, vfc2_w, vfc2_b;
    size_t head_w;   // end of the host blob; device-only [3][C] copy of the 1x1 head-conv weights behind it
    size_t total;    // device allocation
};
ResNetLayout resnet_layout(const oz_resnet_shape &s);

// Board counts are padded to this inside the evaluator: a CTA pair convolves two 16-board units at a time.
constexpr int kBoardPad = 32;
// Pass counters of the dynamic scheduler: one per convolution launch of a forward (1 + 2 * blocks <= this).
constexpr int kMaxConvLayers = 128;
bool conv_static_schedule();
// Cluster tower for launch-bound batches: 1 on, 0 off (per-layer launches), -1 back to the environment (OZ_TOWER, default on).
void set_tower_mode(int mode);
void set_heads_boards(int boards);
long long *tower_dbg_ptr();   // timeline of the last tower launch (OZ_TOWER_DBG set), or null
int tower_max_clusters();      // 16-CTA clusters of the tower the current device runs at once (0: tower unusable)

class ResNet {
  public:
    ~ResNet();
    int load(const oz_resnet_shape &shape, const void *blob, size_t bytes);
    bool loaded() const { return d_blob_ != nullptr; }
    const oz_resnet_shape &shape() const { return shape_; }
    // Capacity in boards (rounded up to a multiple of 32) of workspace `ws`; (re)allocates activations + tensor maps.
    // Workspaces are independent activation sets, so forward() calls on different workspaces may run concurrently on
    // different streams (the engine's pipelined search: one workspace per game group).
    int reserve(size_t boards, int ws = 0);
    // The network's own input tensor of workspace `ws`: bf16 NHWC [boards padded to 32][64][32] (channel c = plane c,
    // channels >= in_planes zero).  A caller that writes it directly passes d_features = nullptr to forward().
    int input_buffer(size_t boards, int ws, void **x0);
    // features: DEVICE float [n][in_planes][64] (reference plane format), or nullptr (see input_buffer).  policy [n][65], value [n].
    // conv_begin / conv_end (optional) are recorded around the convolution launches.
    int forward(const float *d_features, size_t n, float *d_policy, float *d_value, cudaStream_t stream,
                cudaEvent_t conv_begin = nullptr, cudaEvent_t conv_end = nullptr, int ws = 0);
    // Chain the convolution layers with programmatic dependent launch (default on).  Off for concurrent forwards: a
    // layer that is resident early and waits for its predecessor would keep the other stream's CTAs off the SM.
    void set_pdl(bool on) { pdl_ = on; }
    // Kernel launches one forward() issues.
    int launches_per_forward() const { return 2 + (1 + 2 * shape_.blocks); }
    // Launch-bound batches of this network run on the cluster tower (128 channels, the device can launch 16-CTA clusters).
    bool tower_capable() const;
    // Debug hook: one 3x3 convolution layer on device NHWC bf16 buffers already padded to 32 boards.
    // d_work: one device int for the dynamic pass scheduler (zeroed here, in stream order), or null = static schedule.
    static int conv3x3(const void *d_in, int cin, const void *d_w, const float *d_bias, const void *d_skip,
                       int relu, void *d_out, int cout, size_t boards_padded, cudaStream_t stream, int *d_work = nullptr,
                       float *d_stats = nullptr, int stats_copies = 1);
    // d_stats (training forward): [stats_copies][2 * cout] accumulators (zeroed by the caller) that receive the per-channel
    // sum and sum of squares of the fp32 convolution results (BatchNorm batch statistics); requires d_skip == nullptr.

  private:
    oz_resnet_shape shape_{};
    ResNetLayout layout_;
    char *d_blob_ = nullptr;
    struct Workspace {
        size_t cap = 0;
        void *x0 = nullptr, *a = nullptr, *t = nullptr;       // bf16 NHWC activations
        float *head_pre = nullptr;                            // [boards][64][4] fused head-conv pre-activations
        int *work = nullptr;                                  // [kMaxConvLayers] pass counters (dynamic pass scheduler)
        CUtensorMap map_x0, map_a, map_t;                     // 4-row boxes: two board rows per pass
        CUtensorMap map_x0_1, map_a_1, map_t_1;               // 3-row boxes: one board row per pass (small batches)
    };
    void free_workspaces();
    int launch_heads(const Workspace &w, size_t n, bool fuse_heads, float *d_policy, float *d_value, cudaStream_t stream, bool pdl = false);
    int launch_tower(const Workspace &w, size_t pad, cudaStream_t stream);   // every conv layer in one cluster launch (<= 32 boards per cluster)
    std::vector<Workspace> ws_;
    bool pdl_ = true;
    bool dynamic_ = true;                                     // dynamic pass scheduler in the convolutions (OZ_CONV_STATIC=1: off)
    std::vector<CUtensorMap> map_w_;
    CUtensorMap map_wr_{};                                    // residual-layer weights as one 3-D map (cluster tower)
    bool has_wr_ = false;
};

// TMA map of an NHWC bf16 tensor [boards][8][8][C] viewed as (channel, board, x, y) with a box of (32 channels, 16 boards,
// box_x, box_y), SWIZZLE_64B, out-of-bounds coordinates (the halo) filled with zeros.  In shared memory a box is rows of
// 64 bytes, row = ((y * box_x + x) * 16 + board).
int make_act_map(CUtensorMap *m, void *ptr, size_t boards, int C, int box_y, int box_x = 9);

// Weight gradient of a 3x3 "same" convolution on tcgen05 (oz_wgrad.cu): G[co][tap][ci] (fp32, row stride cinp) =
// sum over positions p of dZ[p][co] * X[p + tap][ci]; X NHWC bf16 [boards][8][8][cinp] (cinp = 32 or a multiple of 128),
// dZ NHWC bf16 [boards][8][8][cout] (cout a multiple of 128), boards a multiple of 16.  G is zeroed here (in stream order).
int launch_wgrad3x3(const void *d_x, int cinp, const void *d_dz, int cout, size_t boards, float *d_G, cudaStream_t stream);

// float planes [n][planes][64] -> bf16 NHWC [n_pad][64][32] (channels >= planes and boards >= n are zero).
void launch_planes_to_nhwc32(const float *d_features, int planes, size_t n, size_t n_pad, void *d_out, cudaStream_t stream);

}  // namespace oz
