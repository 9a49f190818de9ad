// The reference's own binding layer over the B200 engine: a pybind11 extension with the module name, classes, methods,
// keyword order and defaults of cpp/src/lib/othello_mcts.cpp:49-151, whose bodies call the C ABI of
// include/othello_b200.h (libothello_b200.so) instead of the reference's C++ MCTS.  LibTorch is used for exactly what
// the reference uses it for at this boundary -- the tensors handed to / taken from Python -- plus reading the
// AlphaZeroNet's parameters (tensors the Python side owns) to fold them into the native evaluator's weight blob.
//
// Status codes are turned back into the exceptions pybind11 made of the reference's std::invalid_argument /
// std::out_of_range (mcts.cpp:167-241, position.h:285-427), with the same message texts.
//
// Built by othello-alphazero_b200/build_ext.py (g++, pybind11, LibTorch headers) into
// othello-alphazero_b200/cpp_binding/othello_mcts/_othello_mcts_impl*.so.
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>
#include <torch/extension.h>

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "othello_b200.h"

namespace py = pybind11;
using namespace py::literals;

namespace {

void check(int rc) {                                   // oz_status -> the reference's exception types
    if (rc == OZ_OK) return;
    const std::string msg = oz_last_error();
    if (rc == OZ_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);   // -> ValueError
    if (rc == OZ_ERR_OUT_OF_RANGE) throw std::out_of_range(msg);           // -> IndexError
    throw std::runtime_error(msg);                                         // -> RuntimeError
}

// ---------------------------------------------------------------- Position (othello_mcts.cpp:68-87, position.h:39-149)
struct Position {
    oz_position p{};
    static Position initial_position() { Position r; check(oz_position_initial(&r.p)); return r; }
    int player() const { return p.player; }
    std::uint64_t player1_discs() const { return p.p1; }
    std::uint64_t player2_discs() const { return p.p2; }
    std::uint64_t legal_moves() const { return p.legal; }
    int at(int index) const { int s = 0; check(oz_position_at(&p, index, &s)); return s; }
    bool is_legal_move(int index) const { int l = 0; check(oz_position_is_legal_move(&p, index, &l)); return l != 0; }
    std::vector<int> legal_actions() const {
        int a[40], n = 0;
        check(oz_position_legal_actions(&p, a, &n));
        return std::vector<int>(a, a + n);
    }
    Position apply_move(std::uint64_t move_mask) const { Position r; check(oz_position_apply_move(&p, move_mask, &r.p)); return r; }
    Position apply_pass() const { Position r; check(oz_position_apply_pass(&p, &r.p)); return r; }
    Position apply_action(int action) const { Position r; check(oz_position_apply_action(&p, action, &r.p)); return r; }
    bool is_terminal() const { return p.player == 0; }
    std::string to_string() const { char buf[512]; check(oz_position_to_string(&p, buf, sizeof buf)); return buf; }
};

// ---------------------------------------------------------------- AlphaZeroNet -> weight blob (neural_net.py:138-172)
py::object unwrap(py::object net) {                    // torch.compile's OptimizedModule (train.py:252-258)
    return py::hasattr(net, "_orig_mod") ? py::object(net.attr("_orig_mod")) : net;
}

bool is_alphazero_net(const py::object &net) {
    for (const char *a : {"conv_block", "residual_blocks", "policy_head", "value_head", "parameters", "buffers"})
        if (!py::hasattr(net, a)) return false;
    return true;
}

torch::Tensor tensor_of(const py::object &module, const char *name) { return module.attr(name).cast<torch::Tensor>().detach(); }

// Eval-mode BatchNorm folded into the preceding convolution: W' = W * s, b' = (b - mean) * s + beta, s = gamma / sqrt(var + eps).
std::pair<torch::Tensor, torch::Tensor> fold(const py::object &conv, const py::object &bn) {
    const auto f64 = torch::TensorOptions().dtype(torch::kFloat64).device(torch::kCPU);
    torch::Tensor w = tensor_of(conv, "weight").to(f64);
    torch::Tensor b = conv.attr("bias").is_none() ? torch::zeros({w.size(0)}, f64) : tensor_of(conv, "bias").to(f64);
    const double eps = bn.attr("eps").cast<double>();
    torch::Tensor s = tensor_of(bn, "weight").to(f64) / torch::sqrt(tensor_of(bn, "running_var").to(f64) + eps);
    w = w * s.view({-1, 1, 1, 1});
    b = (b - tensor_of(bn, "running_mean").to(f64)) * s + tensor_of(bn, "bias").to(f64);
    return {w.to(torch::kFloat32), b.to(torch::kFloat32)};
}

void put(std::vector<char> &blob, int64_t off, const torch::Tensor &t) {
    torch::Tensor c = t.contiguous();
    std::memcpy(blob.data() + off, c.data_ptr(), (size_t)c.numel() * c.element_size());
}

// [cout][cin][3][3] f32 -> tap-major bf16 [ky*3+kx][cout][cin_pad]
torch::Tensor conv3x3_taps(const torch::Tensor &w, int64_t cin_pad) {
    const int64_t cout = w.size(0), cin = w.size(1);
    torch::Tensor t = torch::zeros({9, cout, cin_pad}, torch::kFloat32);
    t.slice(2, 0, cin).copy_(w.permute({2, 3, 0, 1}).reshape({9, cout, cin}));
    return t.to(torch::kBFloat16);
}

struct Folded {
    oz_resnet_shape shape{};
    std::vector<char> blob;
};

bool shape_of(const py::object &net, oz_resnet_shape *s) {
    try {
        py::object conv = net.attr("conv_block").attr("conv");
        s->in_planes = conv.attr("in_channels").cast<int>();
        s->channels = conv.attr("out_channels").cast<int>();
        s->blocks = (int)py::len(net.attr("residual_blocks"));
        s->value_hidden = net.attr("value_head").attr("linear1").attr("out_features").cast<int>();
        py::object ph = net.attr("policy_head"), vh = net.attr("value_head");
        return (s->channels == 128 || s->channels == 256 || s->channels == 384) && s->in_planes >= 1 && s->in_planes <= 32 &&
               s->value_hidden >= 1 && s->value_hidden <= 512 && ph.attr("conv").attr("out_channels").cast<int>() == 2 &&
               ph.attr("linear").attr("in_features").cast<int>() == 128 && ph.attr("linear").attr("out_features").cast<int>() == 65 &&
               vh.attr("conv").attr("out_channels").cast<int>() == 1 && vh.attr("linear1").attr("in_features").cast<int>() == 64 &&
               !net.attr("training").cast<bool>();
    } catch (const py::error_already_set &) {
        return false;
    }
}

Folded fold_network(const py::object &net, const oz_resnet_shape &shape) {
    Folded f;
    f.shape = shape;
    f.blob.assign(oz_resnet_blob_bytes(&shape), 0);
    std::vector<int64_t> offs(8 * (size_t)shape.blocks + 64);
    int n = 0;
    check(oz_resnet_blob_offsets(&shape, offs.data(), (int)offs.size(), &n));
    int k = 0;
    auto w_b = fold(net.attr("conv_block").attr("conv"), net.attr("conv_block").attr("norm"));
    put(f.blob, offs[k++], conv3x3_taps(w_b.first, 32));
    put(f.blob, offs[k++], w_b.second);
    for (py::handle blk : net.attr("residual_blocks")) {
        for (auto names : {std::make_pair("conv1", "norm1"), std::make_pair("conv2", "norm2")}) {
            auto wb = fold(blk.attr(names.first), blk.attr(names.second));
            put(f.blob, offs[k++], conv3x3_taps(wb.first, shape.channels));
            put(f.blob, offs[k++], wb.second);
        }
    }
    py::object ph = net.attr("policy_head"), vh = net.attr("value_head");
    auto pw = fold(ph.attr("conv"), ph.attr("norm"));
    put(f.blob, offs[k++], pw.first.reshape({2, shape.channels}));
    put(f.blob, offs[k++], pw.second);
    put(f.blob, offs[k++], tensor_of(ph.attr("linear"), "weight").to(torch::kCPU, torch::kFloat32));
    put(f.blob, offs[k++], tensor_of(ph.attr("linear"), "bias").to(torch::kCPU, torch::kFloat32));
    auto vw = fold(vh.attr("conv"), vh.attr("norm"));
    put(f.blob, offs[k++], vw.first.reshape({shape.channels}));
    put(f.blob, offs[k++], vw.second);
    put(f.blob, offs[k++], tensor_of(vh.attr("linear1"), "weight").to(torch::kCPU, torch::kFloat32));
    put(f.blob, offs[k++], tensor_of(vh.attr("linear1"), "bias").to(torch::kCPU, torch::kFloat32));
    put(f.blob, offs[k++], tensor_of(vh.attr("linear2"), "weight").to(torch::kCPU, torch::kFloat32).reshape({-1}));
    put(f.blob, offs[k++], tensor_of(vh.attr("linear2"), "bias").to(torch::kCPU, torch::kFloat32));
    if (k != n) throw std::runtime_error("weight blob layout mismatch");
    return f;
}

// Changes whenever a parameter or buffer is updated in place or replaced (re-fold trigger).
std::vector<std::pair<std::uintptr_t, std::int64_t>> fingerprint(const py::object &net) {
    std::vector<std::pair<std::uintptr_t, std::int64_t>> fp;
    for (const char *what : {"parameters", "buffers"})
        for (py::handle h : net.attr(what)()) {
            torch::Tensor t = h.cast<torch::Tensor>();
            fp.emplace_back((std::uintptr_t)t.data_ptr(), (std::int64_t)t._version());
        }
    return fp;
}

// ---------------------------------------------------------------- MCTS (othello_mcts.cpp:89-150, mcts.h:35-216)
class MCTS {
  public:
    MCTS(int history_size, const std::string &torch_device, bool torch_pin_memory, int num_simulations, int num_threads,
         int batch_size, float c_puct_base, float c_puct_init, float dirichlet_epsilon, float dirichlet_alpha,
         std::uint64_t seed, std::uint32_t game_id, int sym_mode)
        : device_(torch_device), pin_(torch_pin_memory) {
        oz_config c;
        oz_config_default(&c);
        c.num_games = 1;
        c.history_size = history_size; c.num_simulations = num_simulations; c.num_threads = num_threads; c.batch_size = batch_size;
        c.c_puct_base = c_puct_base; c.c_puct_init = c_puct_init; c.dirichlet_epsilon = dirichlet_epsilon; c.dirichlet_alpha = dirichlet_alpha;
        c.seed = seed; c.first_game_id = game_id; c.sym_mode = sym_mode; c.evaluator = OZ_EVAL_EXTERNAL;
        c.device = cuda_index();
        check(oz_engine_create(&c, &e_));
    }
    ~MCTS() { oz_engine_destroy(e_); }
    MCTS(const MCTS &) = delete;
    MCTS &operator=(const MCTS &) = delete;

    void reset_position() { check(oz_engine_reset(e_)); }
    Position position() { Position r; check(oz_engine_positions(e_, &r.p)); return r; }

    void search(py::object neural_net) {                       // mcts.h:220-256 + othello_mcts.cpp:36-45, 115-120
        if (py::isinstance<py::str>(neural_net)) {              // built-in evaluators (tests): "synth" / "hash"
            const std::string name = neural_net.cast<std::string>();
            if (name != "synth" && name != "hash") throw std::invalid_argument("unknown built-in evaluator '" + name + "'");
            set_evaluator(name == "synth" ? OZ_EVAL_SYNTH : OZ_EVAL_HASH);
            check(oz_engine_search(e_));
            return;
        }
        py::object net = unwrap(neural_net);
        oz_resnet_shape shape{};
        if (is_alphazero_net(net) && shape_of(net, &shape) && shape.in_planes == 1 + 2 * config().history_size) {
            auto fp = fingerprint(net);
            if (fp != fingerprint_) {                           // weights changed since the last upload: fold again
                Folded f = fold_network(net, shape);
                check(oz_engine_load_resnet(e_, &f.shape, f.blob.data(), f.blob.size()));
                fingerprint_ = std::move(fp);
            }
            set_evaluator(OZ_EVAL_RESNET);
            check(oz_engine_search(e_));                        // tree kernels + tcgen05 evaluator, all on the device
            return;
        }
        search_with_callable(neural_net);
    }

    std::vector<int> visit_counts() {                           // mcts.cpp:45-52
        int32_t n = 0, c[OZ_MAX_CHILDREN];
        check(oz_engine_root_stats(e_, &n, c, nullptr, nullptr));
        return std::vector<int>(c, c + n);
    }
    std::vector<float> mean_action_values() {                   // mcts.cpp:54-61
        int32_t n = 0;
        float q[OZ_MAX_CHILDREN];
        check(oz_engine_root_stats(e_, &n, nullptr, q, nullptr));
        return std::vector<float>(q, q + n);
    }
    py::dict self_play_data() {                                 // mcts.cpp:63-112: lists of 8 CPU tensors
        const int planes = 1 + 2 * config().history_size;
        torch::Tensor f = torch::empty({8, planes, 8, 8}, torch::kFloat32), p = torch::empty({8, 65}, torch::kFloat32);
        check(oz_engine_self_play_data(e_, 0, f.data_ptr<float>(), p.data_ptr<float>()));
        std::vector<torch::Tensor> fl, pl;
        for (int t = 0; t < 8; ++t) { fl.push_back(f[t].clone()); pl.push_back(p[t].clone()); }
        return py::dict("features"_a = fl, "policy"_a = pl);
    }
    void apply_action(int action) { int32_t a = action; check(oz_engine_apply_actions(e_, &a)); }

    // getter / setter pairs, othello_mcts.cpp:131-150 (same validation texts: mcts.cpp:167-241)
    int history_size() { return config().history_size; }
    void set_history_size(int v) { check(oz_engine_set_history_size(e_, v)); }
    std::string torch_device() const { return device_; }
    void set_torch_device(const std::string &v) { device_ = v; }
    bool torch_pin_memory() const { return pin_; }
    void set_torch_pin_memory(bool v) { pin_ = v; }
    int num_simulations() { return config().num_simulations; }
    void set_num_simulations(int v) { check(oz_engine_set_num_simulations(e_, v)); }
    int num_threads() { return config().num_threads; }
    void set_num_threads(int v) { check(oz_engine_set_num_threads(e_, v)); }
    int batch_size() { return config().batch_size; }
    void set_batch_size(int v) { check(oz_engine_set_batch_size(e_, v)); }
    float c_puct_base() { return config().c_puct_base; }
    void set_c_puct_base(float v) { check(oz_engine_set_c_puct_base(e_, v)); }
    float c_puct_init() { return config().c_puct_init; }
    void set_c_puct_init(float v) { check(oz_engine_set_c_puct_init(e_, v)); }
    float dirichlet_epsilon() { return config().dirichlet_epsilon; }
    void set_dirichlet_epsilon(float v) { check(oz_engine_set_dirichlet_epsilon(e_, v)); }
    float dirichlet_alpha() { return config().dirichlet_alpha; }
    void set_dirichlet_alpha(float v) { check(oz_engine_set_dirichlet_alpha(e_, v)); }

    // additive: the engine's counters and the Dirichlet-draw injection hook of the tests
    py::dict counters() {
        oz_counters c;
        check(oz_engine_counters(e_, &c));
        return py::dict("simulations"_a = c.simulations, "evaluations"_a = c.evaluations, "nodes_expanded"_a = c.nodes_expanded,
                        "searches"_a = c.searches, "plies"_a = c.plies, "kernel_launches"_a = c.kernel_launches);
    }
    void set_gamma_table(py::object table) {
        if (table.is_none()) { check(oz_engine_set_gamma_table(e_, nullptr, 0)); return; }
        std::vector<float> t = table.cast<std::vector<float>>();
        check(oz_engine_set_gamma_table(e_, t.data(), (int)t.size()));
    }
    int evaluator() const { return evaluator_; }

  private:
    oz_config config() { oz_config c; check(oz_engine_get_config(e_, &c)); return c; }
    void set_evaluator(int kind) {
        if (kind != evaluator_) { check(oz_engine_set_evaluator(e_, kind)); evaluator_ = kind; }
    }
    int cuda_index() const {                                    // "cuda:1" -> 1; "cpu" / "cuda" -> 0 (trees always live on a GPU)
        const auto pos = device_.find(':');
        if (device_.rfind("cuda", 0) == 0 && pos != std::string::npos) return std::stoi(device_.substr(pos + 1));
        return 0;
    }

    // The evaluator contract of othello_mcts.cpp:36-45: features [B, 1+2H, 8, 8] float32 on torch_device ->
    // {"policy": [B, 65], "value": [B]}, called on this thread with the GIL held; B = num_threads * batch_size leaves.
    void search_with_callable(py::object neural_net) {
        set_evaluator(OZ_EVAL_EXTERNAL);
        const oz_config c = config();
        const int64_t n = (int64_t)c.num_threads * c.batch_size, planes = 1 + 2 * c.history_size;
        const torch::Device cuda(torch::kCUDA, (c10::DeviceIndex)c.device);
        const torch::Device target(device_);
        const auto opt = torch::TensorOptions().dtype(torch::kFloat32).device(cuda);
        py::object torch_cuda = py::module_::import("torch").attr("cuda");
        py::object guard = torch_cuda.attr("device")(c.device);
        guard.attr("__enter__")();
        struct Exit { py::object g; ~Exit() { g.attr("__exit__")(py::none(), py::none(), py::none()); } } exit_guard{guard};
        torch::Tensor feats = torch::zeros({n, planes, 8, 8}, opt);
        void *stream = reinterpret_cast<void *>(torch_cuda.attr("current_stream")(c.device).attr("cuda_stream").cast<std::uintptr_t>());
        check(oz_engine_search_begin(e_));
        try {
            for (int r = 0, R = oz_engine_rounds(e_); r < R; ++r) {
                check(oz_engine_select_device(e_, feats.data_ptr<float>(), stream));
                py::dict out = neural_net(py::cast(target.is_cuda() ? feats : feats.to(target)));
                torch::Tensor policy = out["policy"].cast<torch::Tensor>().detach().to(opt).contiguous();
                torch::Tensor value = out["value"].cast<torch::Tensor>().detach().to(opt).contiguous();
                if (policy.dim() != 2 || policy.size(0) != n || policy.size(1) != 65 || value.dim() != 1 || value.size(0) != n)
                    throw std::invalid_argument("evaluator returned tensors of the wrong shape, expected (" + std::to_string(n) + ", 65) / (" + std::to_string(n) + ",)");
                check(oz_engine_expand_device(e_, policy.data_ptr<float>(), value.data_ptr<float>(), stream));
            }
        } catch (...) {
            oz_engine_search_end(e_);
            throw;
        }
        check(oz_engine_search_end(e_));
    }

    oz_engine *e_ = nullptr;
    std::string device_;
    bool pin_ = false;
    int evaluator_ = OZ_EVAL_EXTERNAL;
    std::vector<std::pair<std::uintptr_t, std::int64_t>> fingerprint_;
};

}  // namespace

PYBIND11_MODULE(_othello_mcts_impl, m) {                        // same names as othello_mcts.cpp:49-151
    m.def("get_legal_moves", &oz_get_legal_moves, "player_discs"_a, "opponent_discs"_a);
    m.def("get_flips", &oz_get_flips, "move_mask"_a, "player_discs"_a, "opponent_discs"_a);
    m.def("abi_version", &oz_abi_version);

    py::class_<Position>(m, "Position")
        .def_static("initial_position", &Position::initial_position)
        .def("player", &Position::player)
        .def("player1_discs", &Position::player1_discs)
        .def("player2_discs", &Position::player2_discs)
        .def("__getitem__", &Position::at)
        .def("legal_moves", &Position::legal_moves)
        .def("is_legal_move", &Position::is_legal_move)
        .def("legal_actions", &Position::legal_actions)
        .def("apply_move", &Position::apply_move)
        .def("apply_pass", &Position::apply_pass)
        .def("apply_action", &Position::apply_action)
        .def("is_terminal", &Position::is_terminal)
        .def("__str__", &Position::to_string);

    py::class_<MCTS>(m, "MCTS")
        .def(py::init<int, const std::string &, bool, int, int, int, float, float, float, float, std::uint64_t, std::uint32_t, int>(),
             "history_size"_a = 4, "torch_device"_a = "cpu", "torch_pin_memory"_a = false, "num_simulations"_a = 800,
             "num_threads"_a = 2, "batch_size"_a = 16, "c_puct_base"_a = 20000.0f, "c_puct_init"_a = 2.5f,
             "dirichlet_epsilon"_a = 0.25f, "dirichlet_alpha"_a = 0.5f, py::kw_only(), "seed"_a = 0, "game_id"_a = 0,
             "sym_mode"_a = (int)OZ_SYM_RANDOM)
        .def("reset_position", &MCTS::reset_position)
        .def("position", &MCTS::position)
        .def("search", &MCTS::search)
        .def("visit_counts", &MCTS::visit_counts)
        .def("mean_action_values", &MCTS::mean_action_values)
        .def("self_play_data", &MCTS::self_play_data)
        .def("apply_action", &MCTS::apply_action)
        .def("history_size", &MCTS::history_size)
        .def("set_history_size", &MCTS::set_history_size)
        .def("torch_device", &MCTS::torch_device)
        .def("set_torch_device", &MCTS::set_torch_device)
        .def("torch_pin_memory", &MCTS::torch_pin_memory)
        .def("set_torch_pin_memory", &MCTS::set_torch_pin_memory)
        .def("num_simulations", &MCTS::num_simulations)
        .def("set_num_simulations", &MCTS::set_num_simulations)
        .def("num_threads", &MCTS::num_threads)
        .def("set_num_threads", &MCTS::set_num_threads)
        .def("batch_size", &MCTS::batch_size)
        .def("set_batch_size", &MCTS::set_batch_size)
        .def("c_puct_base", &MCTS::c_puct_base)
        .def("set_c_puct_base", &MCTS::set_c_puct_base)
        .def("c_puct_init", &MCTS::c_puct_init)
        .def("set_c_puct_init", &MCTS::set_c_puct_init)
        .def("dirichlet_epsilon", &MCTS::dirichlet_epsilon)
        .def("set_dirichlet_epsilon", &MCTS::set_dirichlet_epsilon)
        .def("dirichlet_alpha", &MCTS::dirichlet_alpha)
        .def("set_dirichlet_alpha", &MCTS::set_dirichlet_alpha)
        .def("counters", &MCTS::counters)
        .def("set_gamma_table", &MCTS::set_gamma_table)
        .def("evaluator", &MCTS::evaluator);
}
