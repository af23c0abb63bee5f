// lj_pcg.h -- PCG32 XSH-RR, bit-exact restatement of the reference's pcg.h:16-68 (integer path),
// plus the fp32 uniform the device path draws (pcg.h:49-57, the float overload).
// One stream per camera path, selected by path_stream(pixel * spp + sample): only the 64-bit state is
// kept in the path record, the increment is rebuilt from the ids.  The path index goes through a 64-bit
// finaliser first: PCG streams whose increments differ in a few low bits are affinely related, and the
// samples of one pixel would otherwise sit on neighbouring streams.
#pragma once
#include "lj_common.h"

namespace lj {

constexpr uint64_t kPcgMult = 6364136223846793005ULL;
constexpr uint64_t kPcgDefaultSeed = 0x31e241f862a1fb5eULL;  // pcg.h:33

struct Pcg { uint64_t state, inc; };

LJ_HD uint32_t pcg_next(Pcg &r) {
    uint64_t old = r.state;
    r.state = old * kPcgMult + (r.inc | 1);
    uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t)(old >> 59u);
    return (xorshifted >> rot) | (xorshifted << ((0u - rot) & 31));
}

LJ_HD uint64_t pcg_inc(uint64_t stream_id) { return (stream_id << 1u) | 1u; }

// splitmix64 finaliser (Steele, Lea & Flood 2014): bijective on 64 bits
LJ_HD uint64_t path_stream(uint64_t path_index) {
    uint64_t z = path_index + 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

LJ_HD Pcg pcg_init(uint64_t stream_id, uint64_t seed) {
    Pcg s;
    s.state = 0;
    s.inc = pcg_inc(stream_id);
    pcg_next(s);
    s.state += seed;
    pcg_next(s);
    return s;
}

// [0,1) with 23 random bits (pcg.h:49-57).
LJ_HD float pcg_uniform(Pcg &r) { return u2f((pcg_next(r) >> 9) | 0x3f800000u) - 1.0f; }

}  // namespace lj
