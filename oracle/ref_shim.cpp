// TEST INFRASTRUCTURE ONLY.  C-ABI window onto the UNMODIFIED reference headers, which are
// compiled from where they lie (-I/root/reference/cpp/src/include); no reference source is
// copied into this repository.  Output goes to oracle/_ref/ (git-ignored).  Used to pin
// oracle/oz_oracle.c and to time the reference's own rules code as the CPU baseline.
#include <cstdint>
#include <vector>

#include "position.h"
#include "transformation.h"  // pulls position_iterator.h / search_node.h (torch-free)

using othello::Position;

namespace {
struct PosOut { int32_t player; uint64_t p1, p2, legal, next_legal; };

// Position has no public field constructor; rebuild one by replaying is impossible in
// general, so mirror its layout (position.h:144-148: int + 4 x u64) for the round trip.
static_assert(sizeof(Position) == sizeof(PosOut), "Position layout changed");

Position from_out(const PosOut *p) {
    Position q = Position::initial_position();
    __builtin_memcpy(&q, p, sizeof(PosOut));
    return q;
}
void to_out(const Position &q, PosOut *p) { __builtin_memcpy(p, &q, sizeof(PosOut)); }

uint64_t perft_rec(const Position &p, int depth) {
    if (depth == 0 || p.is_terminal()) return 1;
    if (p.legal_moves() == 0) return perft_rec(p.apply_pass(), depth - 1);
    uint64_t n = 0, moves = p.legal_moves();
    while (moves) {
        uint64_t mv = moves & (~moves + 1);
        moves ^= mv;
        n += perft_rec(p.apply_move(mv), depth - 1);
    }
    return n;
}
}  // namespace

extern "C" {
uint64_t ref_get_legal_moves(uint64_t a, uint64_t b) { return othello::get_legal_moves(a, b); }
uint64_t ref_get_flips(uint64_t m, uint64_t a, uint64_t b) { return othello::get_flips(m, a, b); }
void ref_initial_position(PosOut *o) { to_out(Position::initial_position(), o); }
void ref_apply_move(const PosOut *p, uint64_t m, PosOut *o) { to_out(from_out(p).apply_move(m), o); }
void ref_apply_pass(const PosOut *p, PosOut *o) { to_out(from_out(p).apply_pass(), o); }
void ref_apply_action(const PosOut *p, int a, PosOut *o) { to_out(from_out(p).apply_action(a), o); }
int ref_legal_actions(const PosOut *p, int *out) {
    std::vector<int> v = from_out(p).legal_actions();
    for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
    return (int)v.size();
}
uint64_t ref_perft(const PosOut *p, int depth) { return perft_rec(from_out(p), depth); }
int ref_transform_action(int a, int t) { return othello::transform_action(a, t); }
// history[0] = current position ... history[n-1] = oldest; chained through SearchNode::parent
// so the reference's own SearchNodePositionIterator walks it (position_iterator.h:24-71).
void ref_positions_to_features(const PosOut *hist, int n, float *f, int H, int t) {
    std::vector<othello::SearchNode> nodes;
    nodes.reserve(n);
    for (int i = 0; i < n; ++i) nodes.push_back(othello::SearchNode{from_out(&hist[i])});
    for (int i = 0; i + 1 < n; ++i) nodes[i].parent = &nodes[i + 1];
    othello::positions_to_features(othello::SearchNodePositionIterator(&nodes[0]),
                                   othello::SearchNodePositionIterator::end(), f, H, t);
}
}
