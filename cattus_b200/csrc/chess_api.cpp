// C ABI over the chess rules (include/cattus_b200_chess.h); used by the tests, the tools and a chess host binding.
#include <string>

#include "../../include/cattus_b200.h"
#include "../../include/cattus_b200_chess.h"
#include "chess_rules.hpp"

static thread_local std::string g_chess_error;

static int chess_fail(const std::string& msg) {
    g_chess_error = msg;
    return CATTUS_B200_EINVAL;
}

extern "C" {

int cattus_b200_chess_position(const char* fen, const uint16_t* moves, uint32_t n_moves, cattus_b200_chess_info* out) {
    if (!fen || !out || (n_moves && !moves)) return chess_fail("null argument");
    if (out->struct_size != sizeof(cattus_b200_chess_info)) return chess_fail("cattus_b200_chess_info struct_size mismatch");
    const sp::ChessRules R;
    sp::ChessPos p;
    const std::string err = R.from_fen(fen, p);
    if (!err.empty()) return chess_fail("bad FEN: " + err);
    sp::ChessRules::Move buf[256];
    for (uint32_t i = 0; i < n_moves; ++i) {
        // is_valid_move (core.rs:174-176): the game is still on and the move is legal
        const int n = R.children(p, buf);
        const sp::ChessRules::Move want = sp::ChessRules::real_move(p, moves[i]);  // real -> this view (an involution)
        bool found = false;
        for (int k = 0; k < n; ++k) found |= buf[k] == want;
        if (!found) return chess_fail("move " + std::to_string(i) + " is not legal here");
        p = R.moved(p, want);
    }
    const int n = R.children(p, buf);
    const uint32_t size = out->struct_size;
    std::memset(out, 0, sizeof(*out));
    out->struct_size = size;
    out->turn = p.turn;
    out->status = p.st;
    out->fifty_rule_count = p.fifty;
    out->in_check = p.checkers != 0;
    out->n_legal = static_cast<uint32_t>(n);
    R.planes(p, out->planes);
    for (int k = 0; k < n; ++k) {
        const int idx = R.nn_idx(buf[k]);
        out->moves[k] = sp::ChessRules::real_move(p, buf[k]);
        out->nn_index[k] = static_cast<uint16_t>(idx);
        out->legal_bitmap[idx >> 3] |= static_cast<uint8_t>(1u << (idx & 7));
    }
    return CATTUS_B200_OK;
}

int cattus_b200_chess_perft(const char* fen, uint32_t depth, uint64_t* nodes) {
    if (!fen || !nodes) return chess_fail("null argument");
    const sp::ChessRules R;
    sp::ChessPos p;
    const std::string err = R.from_fen(fen, p);
    if (!err.empty()) return chess_fail("bad FEN: " + err);
    *nodes = R.perft(p, static_cast<int>(depth));
    return CATTUS_B200_OK;
}

int cattus_b200_chess_nn_table(uint16_t* table_out, uint32_t cap) {
    const sp::ChessTables& T = sp::chess_tables();
    const uint32_t n = sizeof(T.nn_index) / sizeof(T.nn_index[0]);
    if (!table_out || cap < n) return chess_fail("table_out needs 64 * 64 + 22 * 4 entries");
    std::memcpy(table_out, T.nn_index, sizeof(T.nn_index));
    return CATTUS_B200_OK;
}

const char* cattus_b200_chess_last_error(void) { return g_chess_error.c_str(); }

}  // extern "C"
