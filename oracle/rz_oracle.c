/*
 * rz_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the reference's Reversi bitboard rules and game state
 * machine, used only by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs to CHECK the CUDA path.  The product
 * path (librz_engine.so) never links, loads or calls anything in this file.
 *
 * Pinning: this file is validated against the reference's own Python
 * (imported from /root/reference/src in the build container) by
 * tests/golden/make_golden.py; the vectors it produced are committed under
 * tests/golden/ and replayed by tests/test_oracle.py, together with the
 * reference's own KATs (test/lib/test_bitboard.py:11-112).
 *
 * Each function cites the reference lines it follows
 * (paths relative to /root/reference/src/reversi_zero/).
 *
 * Bit numbering (lib/bitboard.py:11-17): bit i = square y*8+x, bit 0 = top-left.
 */
#include <stdint.h>
#include <stddef.h>

typedef uint64_t u64;

#define LR_MASK 0x7e7e7e7e7e7e7e7eULL /* lib/bitboard.py:55 */
#define TB_MASK 0x00ffffffffffff00ULL /* lib/bitboard.py:56 */

/* lib/bitboard.py:95-104 -- propagate toward lower bit indices (>>). */
static u64 ray_down(u64 own, u64 enemy, u64 mask, int off) {
    u64 e = enemy & mask;
    u64 t = e & (own >> off);
    for (int i = 0; i < 5; ++i) t |= e & (t >> off);
    return ~(own | enemy) & (t >> off);
}

/* lib/bitboard.py:107-116 -- propagate toward higher bit indices (<<). */
static u64 ray_up(u64 own, u64 enemy, u64 mask, int off) {
    u64 e = enemy & mask;
    u64 t = e & (own << off);
    for (int i = 0; i < 5; ++i) t |= e & (t << off);
    return ~(own | enemy) & (t << off);
}

/* lib/bitboard.py:53-67 find_correct_moves */
u64 rzo_find_correct_moves(u64 own, u64 enemy) {
    const u64 both = LR_MASK & TB_MASK;
    u64 m = 0;
    m |= ray_down(own, enemy, LR_MASK, 1);
    m |= ray_down(own, enemy, both, 9);
    m |= ray_down(own, enemy, TB_MASK, 8);
    m |= ray_down(own, enemy, both, 7);
    m |= ray_up(own, enemy, LR_MASK, 1);
    m |= ray_up(own, enemy, both, 9);
    m |= ray_up(own, enemy, TB_MASK, 8);
    m |= ray_up(own, enemy, both, 7);
    return m;
}

/* lib/bitboard.py:119-125 flip_vertical (row y <-> row 7-y) */
u64 rzo_flip_vertical(u64 x) {
    const u64 k1 = 0x00FF00FF00FF00FFULL, k2 = 0x0000FFFF0000FFFFULL;
    x = ((x >> 8) & k1) | ((x & k1) << 8);
    x = ((x >> 16) & k2) | ((x & k2) << 16);
    return (x >> 32) | (x << 32);
}

/* lib/bitboard.py:141-151 flip_diag_a1h8 (three delta swaps) */
u64 rzo_flip_diag_a1h8(u64 x) {
    const u64 k1 = 0x5500550055005500ULL, k2 = 0x3333000033330000ULL, k4 = 0x0f0f0f0f00000000ULL;
    u64 t;
    t = k4 & (x ^ (x << 28)); x ^= t ^ (t >> 28);
    t = k2 & (x ^ (x << 14)); x ^= t ^ (t >> 14);
    t = k1 & (x ^ (x << 7));  x ^= t ^ (t >> 7);
    return x;
}

/* lib/bitboard.py:154 rotate90 = diag(flipv(x)) */
u64 rzo_rotate90(u64 x) { return rzo_flip_diag_a1h8(rzo_flip_vertical(x)); }
/* lib/bitboard.py:158 rotate180 */
u64 rzo_rotate180(u64 x) { return rzo_rotate90(rzo_rotate90(x)); }

/* lib/bitboard.py:132 bit_count */
int rzo_bit_count(u64 x) { return __builtin_popcountll(x); }

/* lib/bitboard.py:84-92 _calc_flip_half: 4 rays toward higher bits, carry trick.
 * masks are shifted by pos and truncated to 64 bits (b64()). */
static u64 flip_half(int pos, u64 own, u64 enemy) {
    const u64 el[4] = {enemy, enemy & LR_MASK, enemy & LR_MASK, enemy & LR_MASK};
    const u64 base[4] = {0x0101010101010100ULL, 0x00000000000000feULL,
                         0x0002040810204080ULL, 0x8040201008040200ULL};
    u64 flipped = 0;
    for (int i = 0; i < 4; ++i) {
        u64 mask = base[i] << pos;
        u64 outflank = mask & ((el[i] | ~mask) + 1) & own;
        flipped |= (outflank - (outflank != 0)) & mask;
    }
    return flipped;
}

/* lib/bitboard.py:70-81 calc_flip; no legality / occupancy check on pos. */
u64 rzo_calc_flip(int pos, u64 own, u64 enemy) {
    u64 f1 = flip_half(pos, own, enemy);
    u64 f2 = flip_half(63 - pos, rzo_rotate180(own), rzo_rotate180(enemy));
    return f1 | rzo_rotate180(f2);
}

/* ---- game state machine: env/reversi_env.py:18-130 ---- */
typedef struct {
    u64 black, white;
    uint8_t next_player; /* Player enum value: 1 = black, 2 = white (reversi_env.py:9) */
    uint8_t turn;
    uint8_t done;
    uint8_t winner;      /* 0 = None, Winner enum: 1 black, 2 white, 3 draw (reversi_env.py:11) */
} rzo_env;

/* reversi_env.py:26-32 + Board.__init__ :133-140 */
void rzo_env_reset(rzo_env* e) {
    e->black = (0x10ULL << 24) | (0x08ULL << 32);
    e->white = (0x08ULL << 24) | (0x10ULL << 32);
    e->next_player = 1; e->turn = 0; e->done = 0; e->winner = 0;
}

/* reversi_env.py:34-40 update(); NB Board() maps a 0 bitboard to the start stones (:135-136). */
void rzo_env_update(rzo_env* e, u64 black, u64 white, int next_player) {
    e->black = black ? black : ((0x10ULL << 24) | (0x08ULL << 32));
    e->white = white ? white : ((0x08ULL << 24) | (0x10ULL << 32));
    e->next_player = (uint8_t)next_player;
    e->turn = (uint8_t)(rzo_bit_count(e->black) + rzo_bit_count(e->white) - 4);
    e->done = 0; e->winner = 0;
}

/* reversi_env.py:76-85 _game_over */
static void game_over(rzo_env* e) {
    e->done = 1;
    if (e->winner == 0) {
        int b = rzo_bit_count(e->black), w = rzo_bit_count(e->white);
        e->winner = b > w ? 1 : (b < w ? 2 : 3);
    }
}

/* reversi_env.py:42-74 step(); action < 0 means None (= resign, :49-51). */
void rzo_env_step(rzo_env* e, int action) {
    if (action < 0) { /* _resigned :95-97 -> other player wins */
        e->winner = (e->next_player == 1) ? 2 : 1;
        game_over(e);
        return;
    }
    int black_to_move = (e->next_player == 1);
    u64 own = black_to_move ? e->black : e->white;
    u64 enemy = black_to_move ? e->white : e->black;
    u64 flipped = rzo_calc_flip(action, own, enemy);
    if (flipped == 0) { /* illegal_move_to_lose :90-93 */
        e->winner = black_to_move ? 2 : 1;
        game_over(e);
        return;
    }
    own ^= flipped; own |= 1ULL << action; enemy ^= flipped;
    if (black_to_move) { e->black = own; e->white = enemy; }
    else               { e->white = own; e->black = enemy; }
    e->turn += 1;
    if (rzo_find_correct_moves(enemy, own))      e->next_player = black_to_move ? 2 : 1; /* :67-68 */
    else if (rzo_find_correct_moves(own, enemy)) { /* :69-70 auto-pass: same player again */ }
    else game_over(e);                                                                 /* :71-72 */
}

/* ---- batch wrappers (so numpy can drive the oracle without per-call ctypes cost) ---- */
void rzo_find_correct_moves_batch(const u64* own, const u64* enemy, u64* out, size_t n) {
    for (size_t i = 0; i < n; ++i) out[i] = rzo_find_correct_moves(own[i], enemy[i]);
}
void rzo_calc_flip_batch(const uint8_t* pos, const u64* own, const u64* enemy, u64* out, size_t n) {
    for (size_t i = 0; i < n; ++i) out[i] = rzo_calc_flip(pos[i], own[i], enemy[i]);
}
/* SoA step over n independent envs; action -1 = resign. */
void rzo_step_batch(u64* black, u64* white, uint8_t* next_player, uint8_t* turn, uint8_t* done,
                    uint8_t* winner, const int8_t* action, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        rzo_env e = {black[i], white[i], next_player[i], turn[i], done[i], winner[i]};
        rzo_env_step(&e, action[i]);
        black[i] = e.black; white[i] = e.white; next_player[i] = e.next_player;
        turn[i] = e.turn; done[i] = e.done; winner[i] = e.winner;
    }
}
/* transform t in 0..7: bit2 = flip_vertical first, then (t&3) x rotate90 -- the order used by
 * agent/player.py:166-179 (records) and :300-305 (NN input). */
u64 rzo_dihedral(u64 x, int t) {
    if (t & 4) x = rzo_flip_vertical(x);
    for (int i = 0; i < (t & 3); ++i) x = rzo_rotate90(x);
    return x;
}
void rzo_dihedral_batch(const u64* x, const uint8_t* t, u64* out, size_t n) {
    for (size_t i = 0; i < n; ++i) out[i] = rzo_dihedral(x[i], t[i]);
}
