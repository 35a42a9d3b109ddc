// rz_playdata.cu -- host-side writer of the reference's play_data files (worker/self_play.py:180-194,
// lib/data_helper.py:23-25): a JSON array of [[own, enemy], [p0..p63], z] records that the reference's
// optimize worker loads unchanged (worker/optimize.py:215-231).  Text formatting follows CPython's
// json.dump (", " separators, float repr) so files are byte-identical to what the reference would write
// for the same records.
#include <charconv>
#include <string>
#include <stdio.h>
#include <string.h>
#include "rz_bitboard.cuh"
#include "rz_common.cuh"

namespace rz {

// CPython float.__repr__: shortest round-trip digits; fixed notation for -4 <= exp10 < 16, else
// scientific with a sign and at least two exponent digits.
static void append_pyfloat(std::string& out, double v) {
    if (v == 0.0) { out += "0.0"; return; }
    if (v != v) { out += "NaN"; return; }                       // json.dumps(float("nan")) (allow_nan=True, the default)
    if (v - v != 0.0) { out += v > 0 ? "Infinity" : "-Infinity"; return; }
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::scientific);
    *r.ptr = 0;
    // buf = [-]d[.ddd]e[+-]XX
    const char* s = buf;
    if (*s == '-') { out += '-'; ++s; }
    const char* e = strchr(s, 'e');
    std::string digits;
    for (const char* q = s; q < e; ++q) if (*q != '.') digits += *q;
    const int exp10 = atoi(e + 1);
    const int nd = (int)digits.size();
    if (exp10 >= -4 && exp10 < 16) {
        if (exp10 < 0) {
            out += "0.";
            out.append((size_t)(-exp10 - 1), '0');
            out += digits;
        } else if (nd <= exp10 + 1) {
            out += digits;
            out.append((size_t)(exp10 + 1 - nd), '0');
            out += ".0";
        } else {
            out.append(digits, 0, (size_t)exp10 + 1);
            out += '.';
            out.append(digits, (size_t)exp10 + 1, std::string::npos);
        }
    } else {
        out += digits[0];
        if (nd > 1) { out += '.'; out.append(digits, 1, std::string::npos); }
        char eb[16];
        snprintf(eb, sizeof(eb), "e%c%02d", exp10 < 0 ? '-' : '+', exp10 < 0 ? -exp10 : exp10);
        out += eb;
    }
}

static void append_u64(std::string& out, uint64_t v) {
    char buf[32];
    auto r = std::to_chars(buf, buf + sizeof(buf), v);
    out.append(buf, r.ptr);
}

}  // namespace rz

using namespace rz;

extern "C" int rz_write_play_data(const char* path, const rz_game* games, size_t n_games, const rz_ply* plies, int save_policy_of_tau_1,
                                  int change_tau_turn, size_t* n_records) {
    RZ_REQUIRE(path && (n_games == 0 || (games && plies)), "rz_write_play_data: null pointer");
    std::string out;
    out.reserve(n_games * 600000 + 16);
    out += '[';
    size_t nrec = 0;
    for (size_t gi = 0; gi < n_games; ++gi) {
        const rz_game& g = games[gi];
        for (int pid = 1; pid <= 2; ++pid) {  // all of black's records, then all of white's (self_play.py:183)
            const int z = pid == 1 ? g.black_z : -g.black_z;  // player.finish_game(z), self_play.py:230-231
            for (int i = 0; i < g.n_plies; ++i) {
                const rz_ply& pl = plies[g.first_ply + i];
                if (pl.player != pid || !pl.recorded) continue;
                // stored policy, agent/player.py:132,366-385
                double pol[64];
                long long sum = 0;
                int arg = 0;
                for (int a = 0; a < 64; ++a) { sum += pl.n_visit[a]; if (pl.n_visit[a] > pl.n_visit[arg]) arg = a; }
                const int turn = popc64(pl.own) + popc64(pl.enemy) - 4;
                if (save_policy_of_tau_1 || turn < change_tau_turn) {
                    for (int a = 0; a < 64; ++a) pol[a] = (double)pl.n_visit[a] / (double)sum;
                } else {
                    for (int a = 0; a < 64; ++a) pol[a] = 0.0;
                    pol[arg] = 1.0;
                }
                for (int t8 = 0; t8 < 8; ++t8) {  // order: flip in (F,T) x rot_right in 0..3, agent/player.py:166-179
                    const int t = t8;  // t = flip*4 + rot
                    double ps[64];
                    for (int sq = 0; sq < 64; ++sq) ps[dihedral_square(sq, t)] = pol[sq];  // the policy moves with the stones
                    if (nrec) out += ", ";
                    out += "[[";
                    append_u64(out, dihedral(pl.own, t));
                    out += ", ";
                    append_u64(out, dihedral(pl.enemy, t));
                    out += "], [";
                    for (int a = 0; a < 64; ++a) {
                        if (a) out += ", ";
                        append_pyfloat(out, ps[a]);
                    }
                    out += "], ";
                    out += std::to_string(z);
                    out += ']';
                    ++nrec;
                }
            }
        }
    }
    out += ']';
    std::string tmp = std::string(path) + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) { set_error("rz_write_play_data: cannot open %s", tmp.c_str()); return RZ_EIO; }
    const size_t w = fwrite(out.data(), 1, out.size(), f);
    const int cl = fclose(f);
    if (w != out.size() || cl != 0) { set_error("rz_write_play_data: short write to %s", tmp.c_str()); remove(tmp.c_str()); return RZ_EIO; }
    if (rename(tmp.c_str(), path) != 0) { set_error("rz_write_play_data: rename to %s failed", path); remove(tmp.c_str()); return RZ_EIO; }
    if (n_records) *n_records = nrec;
    return RZ_OK;
}
