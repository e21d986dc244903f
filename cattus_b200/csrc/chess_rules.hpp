// Chess rules for the self-play driver (host side of the evaluation path; include/cattus_b200_chess.h exposes them
// for the tests).
//
// Restates engine/src/chess/core.rs (ChessMove::to_idx :55-72, to_nn_idx :93-95 with the table of :453-605,
// ChessPosition::moved_position :326-346, status :348-364, flipped :366-399, equality :292-309) and
// engine/src/chess/net/mod.rs:19-60 (position_to_planes).  The move generator, make-move, en-passant and castling
// bookkeeping of the reference live in the third-party crate `chess` 3.2.0 (engine/Cargo.lock), which is NOT in the
// tree: what is restated here is that crate's published behaviour --
//   * `Board::en_passant()` is the square of the pawn that just advanced two ranks, and only when an enemy pawn stands
//     next to it (`set_ep`);
//   * `MoveGen::new_legal` yields, in this order: pawns, knights, bishops, rooks, queens, king; inside a piece type the
//     pieces that are not pinned in ascending square order, then (when not in check) the pinned ones; en-passant
//     captures after all other pawn moves; per piece the destinations in ascending square order, a promoting
//     destination expanded to queen, knight, rook, bishop; castling is part of the king's destination set;
//   * `Board::status()` is checkmate / stalemate by move count only (no insufficient-material rule).
// The legal move SET is pinned by the published perft counts (tests/test_chess_cpu.py); the ORDER is parity-unpinned
// (restated twice, here and in oracle/chess.py, from the crate's documented structure) and only decides ties in select
// and which child receives which Dirichlet sample.
//
// Representation: a position is ALWAYS stored the way NNetwork::evaluate hands it to the network
// (engine/src/net/mod.rs:158-182): the side to move is "us", plays up the board, and `turn` remembers the real colour.
// ChessPosition::flipped (rank mirror sq ^ 56, colours and castle rights swapped, en-passant file kept) is therefore
// applied once per move, inside moved(), and the evaluator's flip, the planes and the .traindata view cost nothing.
// Moves are kept in the same view; real_move() undoes it (ChessMove::flipped, core.rs:82-91).
#pragma once

#include <cstdint>
#include <cstring>
#include <string>

#include "sp_rules.hpp"

namespace sp {

struct ChessPos {
    uint64_t pc[6];              // pawns, knights, bishops, rooks, queens, kings (both sides)
    uint64_t us;                 // occupancy of the side to move
    uint64_t checkers, pinned;   // enemy pieces giving check; pieces alone between our king and an enemy slider
    uint8_t turn = 1;            // real colour to move: 1 white (Player1), 2 black
    uint8_t castle = 0;          // bit0 us kingside, bit1 us queenside, bit2 them kingside, bit3 them queenside
    uint8_t ep = 64;             // square (this view) of the enemy pawn capturable en passant; 64 = none
    uint8_t fifty = 0;           // fifty_rule_count (core.rs:334-343): white's quiet moves since a pawn move or capture
    uint8_t rev = 0;             // plies since a pawn move or capture: how far back an equal position can lie
    uint8_t st = 0xFF;           // cached status(): 0 ongoing, 1 / 2 winner, 3 draw; 0xFF = not computed yet
    uint8_t pad[2] = {0, 0};
};

enum { kPawn = 0, kKnight, kBishop, kRook, kQueen, kKing };

struct ChessTables {
    uint64_t knight[64], king[64], pawn_up[64];  // pawn_up[s]: the two squares diagonally above s
    uint64_t ray[8][64];                         // N S E W NE NW SE SW, exclusive of the origin
    uint64_t rook_rays[64], bishop_rays[64];
    uint64_t between[64][64], line[64][64];
    uint64_t adjacent_files[8];
    uint16_t nn_index[64 * 64 + 22 * 4];  // MOVE_TO_NN_INDEX (core.rs:597-605); 0xFFFF = not a policy move

    ChessTables() {
        std::memset(this, 0, sizeof(*this));
        static const int dr[8] = {1, -1, 0, 0, 1, 1, -1, -1}, df[8] = {0, 0, 1, -1, 1, -1, 1, -1};
        static const int kn[8][2] = {{1, 2}, {2, 1}, {-1, 2}, {-2, 1}, {1, -2}, {2, -1}, {-1, -2}, {-2, -1}};
        for (int s = 0; s < 64; ++s) {
            const int r = s >> 3, f = s & 7;
            for (auto& d : kn) {
                const int nr = r + d[0], nf = f + d[1];
                if (nr >= 0 && nr < 8 && nf >= 0 && nf < 8) knight[s] |= 1ull << (nr * 8 + nf);
            }
            for (int d = 0; d < 8; ++d) {
                int nr = r + dr[d], nf = f + df[d];
                if (nr >= 0 && nr < 8 && nf >= 0 && nf < 8) king[s] |= 1ull << (nr * 8 + nf);
                while (nr >= 0 && nr < 8 && nf >= 0 && nf < 8) {
                    ray[d][s] |= 1ull << (nr * 8 + nf);
                    nr += dr[d];
                    nf += df[d];
                }
            }
            if (r < 7) {
                if (f > 0) pawn_up[s] |= 1ull << (s + 7);
                if (f < 7) pawn_up[s] |= 1ull << (s + 9);
            }
            rook_rays[s] = ray[0][s] | ray[1][s] | ray[2][s] | ray[3][s];
            bishop_rays[s] = ray[4][s] | ray[5][s] | ray[6][s] | ray[7][s];
        }
        static const int opp[8] = {1, 0, 3, 2, 7, 6, 5, 4};
        for (int a = 0; a < 64; ++a)
            for (int d = 0; d < 8; ++d) {
                uint64_t rest = ray[d][a];
                while (rest) {
                    const int b = ctz64(rest);
                    rest &= rest - 1;
                    between[a][b] = ray[d][a] & ray[opp[d]][b];
                    line[a][b] = ray[d][a] | ray[opp[d]][a] | (1ull << a);
                }
            }
        for (int f = 0; f < 8; ++f) {
            const uint64_t file = 0x0101010101010101ull << f;
            adjacent_files[f] = (f > 0 ? file >> 1 : 0) | (f < 7 ? file << 1 : 0);
        }
        // The 1880-entry policy layout (core.rs:453-595), from its rule: every (from, to) a queen or a knight can
        // travel, from ascending then to ascending; then the rank-7 -> rank-8 promotions by from-file, to-file, q r b n.
        std::memset(nn_index, 0xFF, sizeof(nn_index));
        uint16_t idx = 0;
        for (int f = 0; f < 64; ++f)
            for (int t = 0; t < 64; ++t)
                if (t != f && ((rook_rays[f] | bishop_rays[f] | knight[f]) >> t & 1)) nn_index[f * 64 + t] = idx++;
        for (int sf = 0; sf < 8; ++sf)
            for (int tf = sf - 1; tf <= sf + 1; ++tf)
                if (tf >= 0 && tf < 8)
                    for (int p = 0; p < 4; ++p) nn_index[64 * 64 + (sf * 2 + tf) * 4 + p] = idx++;
    }
};

inline const ChessTables& chess_tables() {
    static const ChessTables t;
    return t;
}

// Move: from | to << 6 | promotion << 12, promotion 0 none, 1 queen, 2 knight, 3 rook, 4 bishop (the order in which the
// crate's move generator yields them).
struct ChessRules {
    using Pos = ChessPos;
    using Move = uint16_t;
    static constexpr bool kChess = true;
    static constexpr int kMaxMoves = 224;  // the serializer's bound is 225 (serialize/chess.rs:34); no position exceeds 218
    static constexpr int kMovesNum = 1880, kLegalBytes = 235, kPlanes = 18;
    const ChessTables& T;
    ChessRules() : T(chess_tables()) {}
    CB2_HD explicit ChessRules(const ChessTables& tables) : T(tables) {}  // device code: the tables live in global memory

    CB2_HD int moves_num() const { return kMovesNum; }
    CB2_HD int max_children() const { return kMaxMoves; }

    CB2_HD static Move make_move(int from, int to, int promo) { return static_cast<Move>(from | (to << 6) | (promo << 12)); }
    CB2_HD static int from_of(Move m) { return m & 63; }
    CB2_HD static int to_of(Move m) { return (m >> 6) & 63; }
    CB2_HD static int promo_of(Move m) { return m >> 12; }
    // ChessMove::flipped (core.rs:82-91)
    CB2_HD static Move flip_move(Move m) { return static_cast<Move>(m ^ (56 | (56 << 6))); }
    CB2_HD static Move real_move(const Pos& p, Move m) { return p.turn == 1 ? m : flip_move(m); }
    // ChessMove::to_idx + MOVE_TO_NN_INDEX (core.rs:55-72, :93-95) of a move in the network's view
    CB2_HD int nn_idx(Move m) const {
        const int from = from_of(m), to = to_of(m), promo = promo_of(m);
        if (promo) {
            const int offset[5] = {0, 0, 3, 1, 2};  // q 0, r 1, b 2, n 3
            return T.nn_index[64 * 64 + ((from & 7) * 2 + (to & 7)) * 4 + offset[promo]];
        }
        return T.nn_index[from * 64 + to];
    }

    CB2_HD static uint64_t occ(const Pos& p) { return p.pc[0] | p.pc[1] | p.pc[2] | p.pc[3] | p.pc[4] | p.pc[5]; }
    CB2_HD static int piece_on(const Pos& p, int sq) {
        for (int k = 0; k < 6; ++k)
            if (p.pc[k] >> sq & 1) return k;
        return -1;
    }
    CB2_HD uint64_t ray_attack(int d, int sq, uint64_t all) const {
        uint64_t r = T.ray[d][sq];
        const uint64_t b = r & all;
        if (b) {
            const bool up = d == 0 || d == 2 || d == 4 || d == 5;  // directions that increase the square index
            const int s = up ? ctz64(b) : 63 - clz64(b);
            r ^= T.ray[d][s];
        }
        return r;
    }
    CB2_HD uint64_t rook_moves(int sq, uint64_t all) const { return ray_attack(0, sq, all) | ray_attack(1, sq, all) | ray_attack(2, sq, all) | ray_attack(3, sq, all); }
    CB2_HD uint64_t bishop_moves(int sq, uint64_t all) const { return ray_attack(4, sq, all) | ray_attack(5, sq, all) | ray_attack(6, sq, all) | ray_attack(7, sq, all); }

    // checkers / pinned of the side to move
    CB2_HD void update_pins(Pos& p) const {
        const uint64_t all = occ(p), them = all & ~p.us;
        const int ksq = ctz64(p.pc[kKing] & p.us);
        p.checkers = 0;
        p.pinned = 0;
        uint64_t pinners = them & ((T.bishop_rays[ksq] & (p.pc[kBishop] | p.pc[kQueen])) | (T.rook_rays[ksq] & (p.pc[kRook] | p.pc[kQueen])));
        while (pinners) {
            const int sq = ctz64(pinners);
            pinners &= pinners - 1;
            const uint64_t bt = T.between[sq][ksq] & all;
            if (bt == 0)
                p.checkers |= 1ull << sq;
            else if ((bt & (bt - 1)) == 0)
                p.pinned |= bt;
        }
        p.checkers |= T.knight[ksq] & them & p.pc[kKnight];
        p.checkers |= T.pawn_up[ksq] & them & p.pc[kPawn];
    }

    CB2_HD Pos initial() const {
        Pos p;
        p.pc[kPawn] = 0x00FF00000000FF00ull;
        p.pc[kKnight] = 0x4200000000000042ull;
        p.pc[kBishop] = 0x2400000000000024ull;
        p.pc[kRook] = 0x8100000000000081ull;
        p.pc[kQueen] = 0x0800000000000008ull;
        p.pc[kKing] = 0x1000000000000010ull;
        p.us = 0xFFFFull;
        p.castle = 15;
        update_pins(p);
        Move buf[256];
        children(p, buf);
        return p;
    }

    // ChessPosition::flipped (core.rs:366-399) on this representation: mirror the ranks and exchange the sides
    CB2_HD void flip_view(Pos& p) const {
        const uint64_t them = occ(p) & ~p.us;
        for (int k = 0; k < 6; ++k) p.pc[k] = bswap64(p.pc[k]);
        p.us = bswap64(them);
        p.castle = static_cast<uint8_t>(((p.castle & 3) << 2) | (p.castle >> 2));
        if (p.ep != 64) p.ep ^= 56;
        p.turn = static_cast<uint8_t>(3 - p.turn);
    }

    CB2_HD bool square_safe_for_king(const Pos& p, int dest) const {  // the crate's legal_king_move
        const uint64_t them = occ(p) & ~p.us;
        const uint64_t all = (occ(p) ^ (p.pc[kKing] & p.us)) | (1ull << dest);
        if (rook_moves(dest, all) & (p.pc[kRook] | p.pc[kQueen]) & them) return false;
        if (bishop_moves(dest, all) & (p.pc[kBishop] | p.pc[kQueen]) & them) return false;
        if (T.knight[dest] & p.pc[kKnight] & them) return false;
        if (T.king[dest] & p.pc[kKing] & them) return false;
        if (T.pawn_up[dest] & p.pc[kPawn] & them) return false;
        return true;
    }
    CB2_HD bool legal_ep(const Pos& p, int src, int dest) const {  // the crate's legal_ep_move
        const uint64_t them = occ(p) & ~p.us;
        const uint64_t all = occ(p) ^ (1ull << p.ep) ^ (1ull << src) ^ (1ull << dest);
        const int ksq = ctz64(p.pc[kKing] & p.us);
        const uint64_t rooks = (p.pc[kRook] | p.pc[kQueen]) & them;
        if ((T.rook_rays[ksq] & rooks) && (rook_moves(ksq, all) & rooks)) return false;
        const uint64_t bishops = (p.pc[kBishop] | p.pc[kQueen]) & them;
        if ((T.bishop_rays[ksq] & bishops) && (bishop_moves(ksq, all) & bishops)) return false;
        return true;
    }

    CB2_HD static int emit(Move* out, int n, int src, uint64_t dests, bool promo) {
        while (dests) {
            const int d = ctz64(dests);
            dests &= dests - 1;
            if (promo) {
                for (int k = 1; k <= 4; ++k) out[n++] = make_move(src, d, k);
            } else {
                out[n++] = make_move(src, d, 0);
            }
        }
        return n;
    }

    // MoveGen::new_legal for the side to move, in the crate's order (see the header comment); returns the count
    CB2_HD int gen(const Pos& p, Move* out) const {
        const uint64_t all = occ(p), us = p.us, mask = ~us;
        const int ksq = ctz64(p.pc[kKing] & us);
        const int n_checkers = popc64(p.checkers);
        int n = 0;
        if (n_checkers <= 1) {
            const bool in_check = n_checkers == 1;
            const uint64_t check_mask = in_check ? (T.between[ctz64(p.checkers)][ksq] ^ p.checkers) : ~0ull;
            const uint64_t pawns = p.pc[kPawn] & us;
            auto pawn_moves = [&](int src) {
                uint64_t m = T.pawn_up[src] & all;
                const uint64_t one = 1ull << (src + 8);
                if (!(all & one)) {
                    m |= one;
                    if ((src >> 3) == 1 && !(all & (one << 8))) m |= one << 8;
                }
                return m & mask;
            };
            for (uint64_t b = pawns & ~p.pinned; b; b &= b - 1) {
                const int src = ctz64(b);
                n = emit(out, n, src, pawn_moves(src) & check_mask, (src >> 3) == 6);
            }
            if (!in_check)
                for (uint64_t b = pawns & p.pinned; b; b &= b - 1) {
                    const int src = ctz64(b);
                    n = emit(out, n, src, pawn_moves(src) & T.line[src][ksq], (src >> 3) == 6);
                }
            if (p.ep != 64) {
                const uint64_t rank = 0xFFull << (p.ep & 56);
                for (uint64_t b = rank & T.adjacent_files[p.ep & 7] & pawns; b; b &= b - 1) {
                    const int src = ctz64(b);
                    if (legal_ep(p, src, p.ep + 8)) out[n++] = make_move(src, p.ep + 8, 0);
                }
            }
            for (int kind = kKnight; kind <= kQueen; ++kind) {
                const uint64_t pieces = p.pc[kind] & us;
                auto pseudo = [&](int src) -> uint64_t {
                    switch (kind) {
                        case kKnight: return T.knight[src] & mask;
                        case kBishop: return bishop_moves(src, all) & mask;
                        case kRook: return rook_moves(src, all) & mask;
                        default: return (bishop_moves(src, all) | rook_moves(src, all)) & mask;
                    }
                };
                for (uint64_t b = pieces & ~p.pinned; b; b &= b - 1) {
                    const int src = ctz64(b);
                    n = emit(out, n, src, pseudo(src) & check_mask, false);
                }
                if (!in_check)
                    for (uint64_t b = pieces & p.pinned; b; b &= b - 1) {
                        const int src = ctz64(b);
                        n = emit(out, n, src, pseudo(src) & T.line[src][ksq], false);
                    }
            }
        }
        uint64_t km = T.king[ksq] & mask;
        for (uint64_t b = km; b; b &= b - 1) {
            const int d = ctz64(b);
            if (!square_safe_for_king(p, d)) km ^= 1ull << d;
        }
        if (n_checkers == 0) {
            if ((p.castle & 1) && !(all & 0x60ull) && square_safe_for_king(p, ksq + 1) && square_safe_for_king(p, ksq + 2)) km ^= 1ull << (ksq + 2);
            if ((p.castle & 2) && !(all & 0x0Eull) && square_safe_for_king(p, ksq - 1) && square_safe_for_king(p, ksq - 2)) km ^= 1ull << (ksq - 2);
        }
        n = emit(out, n, ksq, km, false);
        return n;
    }

    // legal_moves() as the tree sees them (NNetwork::evaluate generates them on the flipped position and maps them
    // back, net/mod.rs:74-87): this view's order.  Also settles status(): ChessPosition::status, core.rs:348-364.
    CB2_HD int children(Pos& p, Move* out) const {
        const int n = gen(p, out);
        if (n == 0)
            p.st = p.checkers ? static_cast<uint8_t>(3 - p.turn) : 3;  // checkmate: the side that moved last wins; stalemate
        else if (p.fifty >= 50)
            p.st = 3;
        else
            p.st = 0;
        return p.st ? 0 : n;
    }
    CB2_HD int status(const Pos& p) const { return p.st; }

    // Board::make_move_new + ChessPosition::moved_position (core.rs:326-346); `m` is in this view; the result is
    // stored in the opponent's view.  status() of the result is settled by children().
    CB2_HD Pos moved(const Pos& p, Move m) const {
        const int src = from_of(m), dst = to_of(m), promo = promo_of(m);
        const uint64_t sb = 1ull << src, db = 1ull << dst;
        const uint64_t all = occ(p), them = all & ~p.us;
        Pos r = p;
        const int piece = piece_on(p, src);
        const bool capture = (all & db) != 0;
        if (capture)
            for (int k = 0; k < 6; ++k) r.pc[k] &= ~db;
        r.pc[piece] &= ~sb;
        const int promo_piece[5] = {0, kQueen, kKnight, kRook, kBishop};
        r.pc[promo ? promo_piece[promo] : piece] |= db;
        r.us = (p.us ^ sb) | db;
        uint8_t castle = p.castle;
        if (src == 4) castle &= static_cast<uint8_t>(~3);
        if (src == 7) castle &= static_cast<uint8_t>(~1);
        if (src == 0) castle &= static_cast<uint8_t>(~2);
        if (dst == 63) castle &= static_cast<uint8_t>(~4);
        if (dst == 56) castle &= static_cast<uint8_t>(~8);
        r.castle = castle;
        r.ep = 64;
        if (piece == kKing && src == 4 && (dst == 6 || dst == 2)) {  // castling: the rook jumps over the king
            const uint64_t rook = dst == 6 ? 0xA0ull : 0x09ull;
            r.pc[kRook] ^= rook;
            r.us ^= rook;
        } else if (piece == kPawn) {
            if (dst - src == 16) {
                if (T.adjacent_files[dst & 7] & (0xFFull << (dst & 56)) & p.pc[kPawn] & them) r.ep = static_cast<uint8_t>(dst);
            } else if (p.ep != 64 && dst == p.ep + 8) {
                r.pc[kPawn] &= ~(1ull << p.ep);
            }
        }
        const bool resets = piece == kPawn || capture;
        r.fifty = resets ? 0 : static_cast<uint8_t>(p.turn == 1 ? p.fifty + 1 : p.fifty);
        r.rev = resets ? 0 : static_cast<uint8_t>(p.rev == 255 ? 255 : p.rev + 1);
        r.st = 0xFF;
        flip_view(r);
        update_pins(r);
        return r;
    }

    // ChessPosition == (core.rs:292-309): boards, castle rights, en passant, side to move -- not the fifty-move count
    CB2_HD static bool same(const Pos& a, const Pos& b) {
        return a.us == b.us && a.turn == b.turn && a.castle == b.castle && a.ep == b.ep && a.pc[0] == b.pc[0] && a.pc[1] == b.pc[1] &&
               a.pc[2] == b.pc[2] && a.pc[3] == b.pc[3] && a.pc[4] == b.pc[4] && a.pc[5] == b.pc[5];
    }

    // position_to_planes (chess/net/mod.rs:19-60) of this view (the side to move plays white)
    CB2_HD void planes(const Pos& p, uint64_t out[18]) const {
        const uint64_t them = occ(p) & ~p.us;
        for (int k = 0; k < 6; ++k) {
            out[k] = p.pc[k] & p.us;
            out[6 + k] = p.pc[k] & them;
        }
        for (int k = 0; k < 4; ++k) out[12 + k] = (p.castle >> k & 1) ? ~0ull : 0ull;
        out[16] = p.ep == 64 ? 0ull : 1ull << p.ep;
        out[17] = ~0ull;
    }

    // Cache key of a position in the evaluator's view: 4 bits per square -- 1..6 our P N B R Q K, 7..12 theirs, 13 / 14
    // our / their rook that may still castle (a castle right implies the rook on its corner and the king on e1/e8),
    // 15 their pawn capturable en passant -- as four bitboards.  Equal keys <=> equal ChessPositions of this view.
    CB2_HD void key_planes(const Pos& p, uint64_t q[4]) const {
        const uint64_t them = occ(p) & ~p.us;
        q[0] = q[1] = q[2] = q[3] = 0;
        for (int k = 0; k < 6; ++k) {
            const uint64_t mine = p.pc[k] & p.us, theirs = p.pc[k] & them;
            const int a = k + 1, b = k + 7;
            for (int j = 0; j < 4; ++j) {
                if (a >> j & 1) q[j] |= mine;
                if (b >> j & 1) q[j] |= theirs;
            }
        }
        uint64_t mine_castle = 0, their_castle = 0;
        if (p.castle & 1) mine_castle |= 1ull << 7;
        if (p.castle & 2) mine_castle |= 1ull << 0;
        if (p.castle & 4) their_castle |= 1ull << 63;
        if (p.castle & 8) their_castle |= 1ull << 56;
        q[0] ^= mine_castle;  // 4 (0100) -> 13 (1101)
        q[3] ^= mine_castle;
        q[2] ^= their_castle;  // 10 (1010) -> 14 (1110)
        if (p.ep != 64) q[3] ^= 1ull << p.ep;  // 7 (0111) -> 15 (1111)
    }

    // ---- test / tooling helpers (include/cattus_b200_chess.h)
    // ChessPosition::from_fen: board, side to move, castle rights and en-passant field (clock fields ignored, as
    // the reference's own fen() omits them).  Returns an empty string on success, else what is wrong.
    std::string from_fen(const char* fen, Pos& out) const {
        Pos p;
        std::memset(p.pc, 0, sizeof(p.pc));
        p.us = p.checkers = p.pinned = 0;
        uint64_t white = 0;
        const char* c = fen;
        int rank = 7, file = 0;
        for (; *c && *c != ' '; ++c) {
            if (*c == '/') {
                if (file != 8) return "bad rank length";
                if (rank == 0 && (c[1] == ' ' || c[1] == 0)) break;  // a trailing '/' (one of the reference's own fixtures has it)
                rank -= 1;
                file = 0;
            } else if (*c >= '1' && *c <= '8') {
                file += *c - '0';
            } else {
                static const char* names = "pnbrqk";
                const char lower = static_cast<char>(*c | 0x20);
                const char* at = std::strchr(names, lower);
                if (!at || rank < 0 || file > 7) return "bad piece placement";
                const int sq = rank * 8 + file;
                p.pc[at - names] |= 1ull << sq;
                if (*c != lower) white |= 1ull << sq;
                file += 1;
            }
        }
        if (rank != 0 || file != 8) return "bad piece placement";
        if (*c == '/') ++c;
        while (*c == ' ') ++c;
        if (*c != 'w' && *c != 'b') return "bad side to move";
        const bool black = *c == 'b';
        ++c;
        while (*c == ' ') ++c;
        uint8_t rights = 0;  // bit0 K, bit1 Q, bit2 k, bit3 q
        for (; *c && *c != ' '; ++c) {
            if (*c == 'K') rights |= 1;
            else if (*c == 'Q') rights |= 2;
            else if (*c == 'k') rights |= 4;
            else if (*c == 'q') rights |= 8;
            else if (*c != '-') return "bad castle field";
        }
        while (*c == ' ') ++c;
        int ep_file = -1;
        if (*c >= 'a' && *c <= 'h') ep_file = *c - 'a';
        const uint64_t all = occ(p);
        if (popc64(p.pc[kKing] & white) != 1 || popc64(p.pc[kKing] & ~white) != 1) return "each side needs exactly one king";
        const uint64_t black_occ = all & ~white;
        auto has = [&](int kind, uint64_t side, int sq) { return (p.pc[kind] & side) >> sq & 1; };
        if (((rights & 3) && !has(kKing, white, 4)) || ((rights & 1) && !has(kRook, white, 7)) || ((rights & 2) && !has(kRook, white, 0)) ||
            ((rights & 12) && !has(kKing, black_occ, 60)) || ((rights & 4) && !has(kRook, black_occ, 63)) || ((rights & 8) && !has(kRook, black_occ, 56)))
            return "castle rights without king and rook on their squares";
        // white's view first, then flip if black is to move
        p.us = white;
        p.turn = 1;
        p.castle = rights;
        p.ep = 64;
        if (black) flip_view(p);
        if (ep_file >= 0) {
            // BoardBuilder: the pawn that just moved stands on ITS fourth rank = rank 5 of this view; Board::set_ep keeps it
            // only if one of our pawns stands next to it
            const int sq = 32 + ep_file;
            const uint64_t them = occ(p) & ~p.us;
            if ((p.pc[kPawn] & them) >> sq & 1)
                if (T.adjacent_files[ep_file] & (0xFFull << 32) & p.pc[kPawn] & p.us) p.ep = static_cast<uint8_t>(sq);
        }
        update_pins(p);
        // the side NOT to move must not be in check
        {
            Pos q = p;
            flip_view(q);
            update_pins(q);
            if (q.checkers) return "the side that just moved is in check";
        }
        Move buf[256];
        children(p, buf);
        out = p;
        return std::string();
    }

    uint64_t perft(const Pos& p, int depth) const {
        Move buf[256];
        const int n = gen(p, buf);
        if (depth <= 1) return depth == 1 ? static_cast<uint64_t>(n) : 1;
        uint64_t total = 0;
        for (int i = 0; i < n; ++i) total += perft(moved(p, buf[i]), depth - 1);
        return total;
    }
};

}  // namespace sp
