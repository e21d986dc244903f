// Hex and TicTacToe rules of the self-play driver, usable from host code (csrc/selfplay.cpp) and from device code (the
// device-resident search of csrc/dsearch_core.hpp): every function the search needs is `CB2_HD`.
//
// Restates engine/src/hex/core.rs:112-335 and engine/src/ttt/core.rs:101-246 of the reference (paths relative to
// /root/reference).  Chess lives in chess_rules.hpp.
#pragma once

#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define CB2_HD __host__ __device__
#else
#define CB2_HD
#endif

namespace sp {

using u128 = unsigned __int128;

// ------------------------------------------------------------------------------------------------ bit helpers
CB2_HD inline int ctz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __ffsll(static_cast<long long>(x)) - 1;
#else
    return __builtin_ctzll(x);
#endif
}
CB2_HD inline int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __clzll(static_cast<long long>(x));
#else
    return __builtin_clzll(x);
#endif
}
CB2_HD inline int popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}
CB2_HD inline uint64_t bswap64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    x = ((x & 0x00FF00FF00FF00FFull) << 8) | ((x >> 8) & 0x00FF00FF00FF00FFull);
    x = ((x & 0x0000FFFF0000FFFFull) << 16) | ((x >> 16) & 0x0000FFFF0000FFFFull);
    return (x << 32) | (x >> 32);
#else
    return __builtin_bswap64(x);
#endif
}
CB2_HD inline int ctz128(u128 x) {
    const uint64_t lo = static_cast<uint64_t>(x);
    return lo ? ctz64(lo) : 64 + ctz64(static_cast<uint64_t>(x >> 64));
}
CB2_HD inline int popcount128(u128 x) { return popc64(static_cast<uint64_t>(x)) + popc64(static_cast<uint64_t>(x >> 64)); }
CB2_HD inline u128 bit128(int i) { return static_cast<u128>(1) << i; }

// Board word W: uint64_t for boards up to 8x8 (every shipped hex config: half the node size and single-instruction bit
// operations), unsigned __int128 up to 11x11 (the reference's u128, hex/core.rs:52-54).
CB2_HD inline int ctz_word(uint64_t x) { return ctz64(x); }
CB2_HD inline int ctz_word(u128 x) { return ctz128(x); }

// status(): 0 ongoing, 1 Player1 won, 2 Player2 won, 3 draw
struct PosKey {
    u128 a, b;
    bool operator==(const PosKey& o) const { return a == o.a && b == o.b; }
};

template <class W>
struct HexPosT {
    W red = 0, blue = 0, left_red_reach = 0, top_blue_reach = 0;
    uint8_t turn = 1, empty = 0, winner = 0;
};

// engine/src/hex/core.rs
template <class W>
struct HexRulesT {
    using Pos = HexPosT<W>;
    using Move = uint8_t;
    using Word = W;
    static constexpr bool kChess = false;
    static constexpr int kMaxChildren = 121;
    CB2_HD int max_children() const { return cells; }
    CB2_HD static W bit(int i) { return static_cast<W>(1) << i; }
    int s = 0, cells = 0;
    W full = 0, col0 = 0, col_last = 0, row0 = 0, row_last = 0;
    W nb[121];
    uint8_t tr[121];

    explicit HexRulesT(int size) : s(size), cells(size * size) {
        full = cells == static_cast<int>(8 * sizeof(W)) ? ~static_cast<W>(0) : static_cast<W>(bit(cells) - 1);
        const int dirs[6][2] = {{0, 1}, {-1, 0}, {-1, -1}, {0, -1}, {1, 0}, {1, 1}};  // core.rs:204
        for (int i = 0; i < 121; ++i) {
            nb[i] = 0;
            tr[i] = 0;
        }
        for (int r = 0; r < s; ++r)
            for (int c = 0; c < s; ++c) {
                const int i = r * s + c;
                tr[i] = static_cast<uint8_t>(c * s + r);
                nb[i] = 0;
                for (auto& d : dirs) {
                    const int nr = r + d[0], nc = c + d[1];
                    if (nr >= 0 && nr < s && nc >= 0 && nc < s) nb[i] |= bit(nr * s + nc);
                }
                if (c == 0) col0 |= bit(i);
                if (c == s - 1) col_last |= bit(i);
                if (r == 0) row0 |= bit(i);
                if (r == s - 1) row_last |= bit(i);
            }
    }
    CB2_HD int moves_num() const { return cells; }
    CB2_HD int words_per_plane() const { return 2; }  // u128 -> [lo, hi] (serialize/hex.rs:16-28); the C ABI uses ceil(S*S/64)
    CB2_HD Pos initial() const {
        Pos p;
        p.empty = static_cast<uint8_t>(cells);
        return p;
    }
    CB2_HD int status(const Pos& p) const {  // core.rs:314-322
        if (p.winner) return p.winner;
        if (p.empty == 0) return 3;
        return 0;
    }
    CB2_HD u128 legal_mask(const Pos& p) const { return static_cast<u128>(static_cast<W>(full & ~(p.red | p.blue))); }  // core.rs:297-305 (ascending index)
    // core.rs:215-264: flood the player's reach map from the new stone; the first end-edge cell reached wins.
    CB2_HD void update_reach(Pos& p, int idx, int player) const {
        const W board = player == 1 ? p.red : p.blue;
        W& reach = player == 1 ? p.left_red_reach : p.top_blue_reach;
        const W begin = player == 1 ? col0 : row0;
        const W end = player == 1 ? col_last : row_last;
        if (!((begin & bit(idx)) || (nb[idx] & reach))) return;
        W layer = bit(idx);
        reach |= layer;
        while (layer) {
            const int i = ctz_word(layer);
            layer &= static_cast<W>(~bit(i));
            if (end & bit(i)) {
                p.winner = static_cast<uint8_t>(player);
            } else {
                const W add = nb[i] & board & static_cast<W>(~reach);
                reach |= add;
                layer |= add;
            }
        }
    }
    CB2_HD Pos moved(const Pos& p, int m) const {  // core.rs:272-285
        Pos r = p;
        if (r.turn == 1)
            r.red |= bit(m);
        else
            r.blue |= bit(m);
        update_reach(r, m, r.turn);
        r.empty -= 1;
        r.turn = static_cast<uint8_t>(3 - r.turn);
        return r;
    }
    CB2_HD W transpose(W bb) const {  // HexBitboard::flip, core.rs:61-71
        W f = 0;
        while (bb) {
            const int i = ctz_word(bb);
            bb &= bb - 1;
            f |= bit(tr[i]);
        }
        return f;
    }
    CB2_HD Pos flipped(const Pos& p) const {  // core.rs:324-334
        Pos r;
        r.red = transpose(p.blue);
        r.blue = transpose(p.red);
        r.turn = static_cast<uint8_t>(3 - p.turn);
        r.left_red_reach = transpose(p.top_blue_reach);
        r.top_blue_reach = transpose(p.left_red_reach);
        r.empty = p.empty;
        r.winner = p.winner ? static_cast<uint8_t>(3 - p.winner) : 0;
        return r;
    }
    // the part of flipped() the evaluator needs (planes, legal mask, cache key): boards and turn only
    CB2_HD Pos flipped_boards(const Pos& p) const {
        Pos r;
        r.red = transpose(p.blue);
        r.blue = transpose(p.red);
        r.turn = static_cast<uint8_t>(3 - p.turn);
        r.empty = p.empty;
        return r;
    }
    CB2_HD int flip_move(int m) const { return tr[m]; }  // core.rs:36-38
    CB2_HD bool same(const Pos& a, const Pos& b) const { return a.red == b.red && a.blue == b.blue && a.turn == b.turn; }
    CB2_HD bool child_matches(const Pos& parent, int m, const Pos& target) const {
        const W red = parent.turn == 1 ? static_cast<W>(parent.red | bit(m)) : parent.red;
        const W blue = parent.turn == 1 ? parent.blue : static_cast<W>(parent.blue | bit(m));
        return red == target.red && blue == target.blue && target.turn == 3 - parent.turn;
    }
    PosKey key(const Pos& p) const { return PosKey{static_cast<u128>(p.red), static_cast<u128>(p.blue)}; }
    // position_to_planes (hex/net.rs:14-24): [red, blue, ones]
    CB2_HD void planes(const Pos& p, u128 out[3]) const {
        out[0] = p.red;
        out[1] = p.blue;
        out[2] = full;
    }
};

struct TttPos {
    uint16_t x = 0, o = 0;
    uint8_t turn = 1, winner = 0;
};

// engine/src/ttt/core.rs
struct TttRules {
    using Pos = TttPos;
    using Move = uint8_t;
    static constexpr bool kChess = false;
    static constexpr int kMaxChildren = 9;
    CB2_HD int max_children() const { return 9; }
    CB2_HD int moves_num() const { return 9; }
    CB2_HD int words_per_plane() const { return 1; }
    CB2_HD Pos initial() const { return Pos(); }
    CB2_HD static uint8_t winner_of(uint16_t x, uint16_t o) {  // core.rs:170-193: x before o on every line, in this order
        const uint16_t lines[8] = {0b111000000, 0b000111000, 0b000000111, 0b100100100, 0b010010010, 0b001001001, 0b100010001, 0b001010100};
        for (int k = 0; k < 8; ++k) {
            const uint16_t w = lines[k];
            if ((x & w) == w) return 1;
            if ((o & w) == w) return 2;
        }
        return 0;
    }
    CB2_HD int status(const Pos& p) const {
        if (p.winner) return p.winner;
        if ((p.x | p.o) == 0x1FF) return 3;
        return 0;
    }
    CB2_HD u128 legal_mask(const Pos& p) const { return static_cast<u128>(0x1FFu & ~(p.x | p.o)); }
    CB2_HD Pos moved(const Pos& p, int m) const {
        Pos r = p;
        if (r.turn == 1)
            r.x |= static_cast<uint16_t>(1u << m);
        else
            r.o |= static_cast<uint16_t>(1u << m);
        r.turn = static_cast<uint8_t>(3 - r.turn);
        r.winner = winner_of(r.x, r.o);
        return r;
    }
    CB2_HD Pos flipped(const Pos& p) const {
        Pos r;
        r.x = p.o;
        r.o = p.x;
        r.turn = static_cast<uint8_t>(3 - p.turn);
        r.winner = p.winner ? static_cast<uint8_t>(3 - p.winner) : 0;
        return r;
    }
    CB2_HD Pos flipped_boards(const Pos& p) const { return flipped(p); }
    CB2_HD int flip_move(int m) const { return m; }
    CB2_HD bool same(const Pos& a, const Pos& b) const { return a.x == b.x && a.o == b.o && a.turn == b.turn; }
    CB2_HD bool child_matches(const Pos& parent, int m, const Pos& target) const {
        const Pos c = moved(parent, m);
        return same(c, target);
    }
    PosKey key(const Pos& p) const { return PosKey{p.x, p.o}; }
    CB2_HD void planes(const Pos& p, u128 out[3]) const {
        out[0] = p.x;
        out[1] = p.o;
        out[2] = 0x1FF;
    }
};

}  // namespace sp
