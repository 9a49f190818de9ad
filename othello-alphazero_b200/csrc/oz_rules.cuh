// Othello bitboard rules shared by the sm_100a kernels and the host-side Position.
//
// Behaviour follows the reference rules header (cpp/src/include/position.h):
//   get_legal_moves  position.h:202-229     get_flips       position.h:231-262
//   apply_move       position.h:328-363     apply_pass      position.h:382-386
//   apply_action     position.h:402-408     initial_position position.h:264-272
// Bit 63 is square a1 (action 0); action i <-> bit (63 - i)  (position.h:275).
//
// The reference walks every direction with a 6-step "dumb" fill.  Here each ray is a
// 4-step prefix-doubling fill (runs of 1, 2, 4, 6 squares: 6 is the longest run a masked
// ray can hold), which gives the same masks bit for bit for every input.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define OZ_HD __host__ __device__ __forceinline__
#else
#define OZ_HD inline
#endif

namespace oz {

constexpr uint64_t kNoLeftRight = 0x7E7E7E7E7E7E7E7Eull;  // position.h:155
constexpr uint64_t kNoTopBottom = 0x00FFFFFFFFFFFF00ull;  // position.h:158
constexpr uint64_t kNoEdges     = kNoLeftRight & kNoTopBottom;

constexpr uint64_t kInitialP1    = 0x0000000810000000ull;
constexpr uint64_t kInitialP2    = 0x0000001008000000ull;
constexpr uint64_t kInitialLegal = 0x0000102004080000ull;

struct Position {
    int32_t  player;      // 1 Black, 2 White, 0 terminal
    uint64_t p1;          // Black discs
    uint64_t p2;          // White discs
    uint64_t legal;       // legal moves of the side to move
    uint64_t next_legal;  // legal moves of the other side when `legal` is empty
};

// Run of `o` squares that starts next to a `gen` square and continues towards higher bits
// (left shift by S): the reference's 6-step fill (position.h:186-196) as a prefix doubling
// 1, 2, 4, 6.  The generator squares themselves are never part of the result, so the two
// agree for every input, including overlapping or otherwise unreachable disc sets.
template <int S>
OZ_HD uint64_t run_up(uint64_t gen, uint64_t o) {
    uint64_t f = o & (gen << S);
    f |= o & (f << S);
    const uint64_t pairs = o & (o << S);
    f |= pairs & (f << (2 * S));
    f |= pairs & (f << (2 * S));
    return f;
}

// Same towards lower bits (right shift by S).
template <int S>
OZ_HD uint64_t run_down(uint64_t gen, uint64_t o) {
    uint64_t f = o & (gen >> S);
    f |= o & (f >> S);
    const uint64_t pairs = o & (o >> S);
    f |= pairs & (f >> (2 * S));
    f |= pairs & (f >> (2 * S));
    return f;
}

OZ_HD uint64_t legal_moves(uint64_t player, uint64_t opp) {
    const uint64_t oh = opp & kNoLeftRight;   // rays along a rank
    const uint64_t ov = opp & kNoTopBottom;   // rays along a file
    const uint64_t od = opp & kNoEdges;       // diagonal rays
    uint64_t m = 0;
    m |= run_up<1>(player, oh) << 1;
    m |= run_down<1>(player, oh) >> 1;
    m |= run_up<8>(player, ov) << 8;
    m |= run_down<8>(player, ov) >> 8;
    m |= run_up<7>(player, od) << 7;
    m |= run_down<7>(player, od) >> 7;
    m |= run_up<9>(player, od) << 9;
    m |= run_down<9>(player, od) >> 9;
    return m & ~(player | opp);
}

OZ_HD uint64_t flips(uint64_t move, uint64_t player, uint64_t opp) {
    const uint64_t oh = opp & kNoLeftRight;
    const uint64_t ov = opp & kNoTopBottom;
    const uint64_t od = opp & kNoEdges;
    uint64_t f = 0, r;
    // A ray counts only when the square after the run holds one of the mover's discs.
    r = run_up<1>(move, oh);   f |= ((r << 1) & player) ? r : 0;
    r = run_down<1>(move, oh); f |= ((r >> 1) & player) ? r : 0;
    r = run_up<8>(move, ov);   f |= ((r << 8) & player) ? r : 0;
    r = run_down<8>(move, ov); f |= ((r >> 8) & player) ? r : 0;
    r = run_up<7>(move, od);   f |= ((r << 7) & player) ? r : 0;
    r = run_down<7>(move, od); f |= ((r >> 7) & player) ? r : 0;
    r = run_up<9>(move, od);   f |= ((r << 9) & player) ? r : 0;
    r = run_down<9>(move, od); f |= ((r >> 9) & player) ? r : 0;
    return f;
}

OZ_HD Position initial_position() {
    Position p;
    p.player = 1; p.p1 = kInitialP1; p.p2 = kInitialP2;
    p.legal = kInitialLegal; p.next_legal = 0;
    return p;
}

OZ_HD Position apply_move(const Position& s, uint64_t move) {
    Position n;
    uint64_t me  = (s.player == 1) ? s.p1 : s.p2;
    uint64_t opp = (s.player == 1) ? s.p2 : s.p1;
    const uint64_t f = flips(move, me, opp);
    me |= move | f;
    opp &= ~f;
    n.player = 3 - s.player;
    n.p1 = (s.player == 1) ? me : opp;
    n.p2 = (s.player == 1) ? opp : me;
    n.legal = legal_moves(opp, me);
    n.next_legal = 0;
    if (n.legal == 0) {
        n.next_legal = legal_moves(me, opp);
        if (n.next_legal == 0) n.player = 0;
    }
    return n;
}

OZ_HD Position apply_pass(const Position& s) {
    Position n;
    n.player = 3 - s.player; n.p1 = s.p1; n.p2 = s.p2;
    n.legal = s.next_legal; n.next_legal = 0;
    return n;
}

OZ_HD Position apply_action(const Position& s, int action) {
    return action == 64 ? apply_pass(s) : apply_move(s, 1ull << (63 - action));
}

// D4 symmetry of an action index (transformation.h:40-57): bit 0 mirrors the columns,
// then (t >> 1) quarter turns (row, col) <- (col, 7 - row).  64 (pass) is a fixed point.
OZ_HD int transform_action(int action, int t) {
    if (action == 64) return 64;
    int row = action >> 3, col = action & 7;
    if (t & 1) col = 7 - col;
    const int rot = t >> 1;
    for (int i = 0; i < rot; ++i) { const int r = row; row = col; col = 7 - r; }
    return row * 8 + col;
}

}  // namespace oz
