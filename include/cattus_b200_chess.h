/* cattus_b200_chess.h -- C ABI over the chess rules the self-play driver uses (cattus_b200/csrc/chess_rules.hpp).
 *
 * The host side of the evaluation path for chess: what a caller needs to turn a position into the evaluator's inputs
 * (18 planes + the 235-byte legal-move bitmap) and its outputs back into per-move probabilities.  Each entry point
 * names the reference function it stands for; the move generator itself is the third-party crate `chess` 3.2.0 in
 * the reference (engine/Cargo.lock), restated here from its published behaviour.
 *
 * Squares are rank * 8 + file with a1 = 0 (engine/src/chess/core.rs:135-138).  A move is
 * from | to << 6 | promotion << 12 with promotion 0 none, 1 queen, 2 knight, 3 rook, 4 bishop, in REAL board
 * coordinates.  A position is given as a FEN (board, side to move, castle rights, en-passant field; clocks ignored --
 * ChessPosition::from_fen, core.rs:170-172) plus a list of moves played from it.
 */
#ifndef CATTUS_B200_CHESS_H
#define CATTUS_B200_CHESS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CATTUS_B200_CHESS_MAX_MOVES 224 /* serialize/chess.rs:34 allows 225; no legal position exceeds 218 */

typedef struct cattus_b200_chess_info {
    uint32_t struct_size;
    uint32_t turn;        /* 1 white (Player1), 2 black: Position::turn, core.rs:318-320 */
    uint32_t status;      /* 0 ongoing, 1 white won, 2 black won, 3 draw: ChessPosition::status, core.rs:348-364 */
    uint32_t fifty_rule_count; /* core.rs:334-343 */
    uint32_t in_check;
    uint32_t n_legal;
    /* The evaluator's inputs for this position (engine/src/net/mod.rs:74-87: flipped so that the side to move plays
     * white): position_to_planes (chess/net/mod.rs:19-60) and the legal-move bitmap over the 1880 nn indices. */
    uint64_t planes[18];
    uint8_t legal_bitmap[235];
    uint8_t pad;
    /* legal_moves() as NNetwork::evaluate returns them: generated on the flipped position, mapped back to real
     * coordinates (net/mod.rs:166-182) -- the order in which MctsPlayer::create_children inserts the children.
     * nn_index[i] is Move::to_nn_idx of the move in the evaluator's view (core.rs:93-95). */
    uint16_t moves[CATTUS_B200_CHESS_MAX_MOVES];
    uint16_t nn_index[CATTUS_B200_CHESS_MAX_MOVES];
} cattus_b200_chess_info;

/* ChessPosition::from_fen(fen) followed by moved_position(m) for each of the n_moves moves (core.rs:326-346).  Fails
 * with CATTUS_B200_EINVAL on a malformed FEN or an illegal move. */
int cattus_b200_chess_position(const char* fen, const uint16_t* moves, uint32_t n_moves, cattus_b200_chess_info* out);

/* Number of move sequences of length `depth` from the position (the usual move-generator check). */
int cattus_b200_chess_perft(const char* fen, uint32_t depth, uint64_t* nodes);

/* MOVE_TO_NN_INDEX (core.rs:597-605): table_out[64 * 64 + 22 * 4], 0xFFFF where there is no policy move. */
int cattus_b200_chess_nn_table(uint16_t* table_out, uint32_t cap);

/* thread-local message of the last failed cattus_b200_chess_* call on this thread */
const char* cattus_b200_chess_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* CATTUS_B200_CHESS_H */
