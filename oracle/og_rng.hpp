// og_rng.hpp -- TEST INFRASTRUCTURE (oracle).  Not part of the product.
//
// The reference draws from OCaml's global `Random` (111 call sites, e.g.
// mcmc.ml:47, interpolate_pdf.ml:115); its stream is version dependent and is
// NOT reproduced.  The oracle and the GPU path instead share one
// *specification* of a counter-based stream, implemented twice (here in plain
// C++, in mcmc_ocaml_b200/csrc/rng.cuh for the device), so that stochastic
// entry points can be compared draw-for-draw on the same seed:
//
//   call key   (k0,k1) = Philox4x32-10(ctr = (epoch_lo, epoch_hi, 'mcmc', 0),
//                                      key = (seed_lo, seed_hi)).words[0..1]
//   block      B(purpose, g, step, blk) = Philox4x32-10(
//                  ctr = (blk, step_lo, g_lo,
//                         g_hi16 | purpose << 16 | step_hi8 << 24), key = call key)
//   draw j of (purpose, g, step): 52 private bits P.  Draws come in groups of 11 from 5 blocks (640 bits for
//                  11 x 52 = 572): with m = j / 11, i = j % 11,
//                    i < 10 : block 5m + i/2, words (a, b) = (w[2(i&1)], w[2(i&1)+1]),
//                             P = (a & 0xFFFFF) << 32 | b          (low 20 bits of the first word : second word)
//                    i = 10 : the 12 TOP bits of the first words that the ten draws above leave unused,
//                             s0 = w0(5m) >> 20, s1 = w2(5m) >> 20, s2 = w0(5m+1) >> 20, s3 = w2(5m+1) >> 20,
//                             s4 = w0(5m+2) >> 20:   P = s0<<40 | s1<<28 | s2<<16 | s3<<4 | s4>>8
//                  (a 10-D proposal + accept test = 11 uniforms = 5 Philox blocks per Metropolis-Hastings step)
//   Random.float 1.0  ->  u52 = bits(0x3FF<<52 | P) - 1.0   in [0,1)
//   Random.int n      ->  mulhi64(P << 12, n)               (every draw uses its own 52 bits only)
//
// Philox4x32-10: Salmon et al., "Parallel random numbers: as easy as 1, 2, 3"
// (SC'11); checked against the Random123 known-answer vectors in
// tests/test_oracle_rng.py.
#pragma once
#include <cstdint>
#include <cstring>

namespace og {

struct Philox {
  static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  static constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  static inline void block(const uint32_t ctr[4], const uint32_t key[2],
                           uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
      uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
      uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0;
      uint32_t h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
      uint32_t n0 = h1 ^ c1 ^ k0, n1 = l1, n2 = h0 ^ c3 ^ k1, n3 = l0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
  }
};

enum Purpose : uint32_t {
  P_MH = 1, P_RJ = 2, P_RJ_INIT = 3, P_DRAW = 4, P_NEST_INIT = 5,
  P_NEST_MCMC = 6, P_NEST_START = 7, P_POST = 8, P_BOOT = 9
};

struct CallKey { uint32_t k[2]; };

inline CallKey derive_key(uint64_t seed, uint64_t epoch) {
  uint32_t ctr[4] = {(uint32_t)epoch, (uint32_t)(epoch >> 32), 0x6d636d63u, 0u};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t out[4];
  Philox::block(ctr, key, out);
  return CallKey{{out[0], out[1]}};
}

// Sequential cursor over the draws of one (purpose, g, step).
struct Rng {
  uint32_t key[2];
  uint32_t ctr[4];
  uint32_t w[4];
  uint32_t j = 0;  // next draw index
  Rng(const CallKey &ck, uint32_t purpose, uint64_t g, uint64_t step) {
    key[0] = ck.k[0]; key[1] = ck.k[1];
    ctr[0] = 0;
    ctr[1] = (uint32_t)step;
    ctr[2] = (uint32_t)g;
    ctr[3] = (uint32_t)((g >> 32) & 0xFFFFu) | ((purpose & 0xFFu) << 16) |
             (uint32_t)(((step >> 32) & 0xFFu) << 24);
  }
  uint32_t sp[5] = {0, 0, 0, 0, 0};  // spare top-12-bit pieces of the current group's first three blocks
  // the 52 private bits of the next draw
  inline uint64_t bits52() {
    const uint32_t m = j / 11u, i = j % 11u;
    uint64_t P;
    if (i < 10u) {
      if ((i & 1u) == 0u) {
        const uint32_t gi = i >> 1;
        ctr[0] = 5u * m + gi; Philox::block(ctr, key, w);
        if (gi == 0u) { sp[0] = w[0] >> 20; sp[1] = w[2] >> 20; }
        else if (gi == 1u) { sp[2] = w[0] >> 20; sp[3] = w[2] >> 20; }
        else if (gi == 2u) { sp[4] = w[0] >> 20; }
      }
      const uint32_t a = w[2 * (i & 1u)], b = w[2 * (i & 1u) + 1];
      P = ((uint64_t)(a & 0xFFFFFu) << 32) | b;
    } else {
      P = ((uint64_t)sp[0] << 40) | ((uint64_t)sp[1] << 28) | ((uint64_t)sp[2] << 16) | ((uint64_t)sp[3] << 4) | (uint64_t)(sp[4] >> 8);
    }
    ++j;
    return P;
  }
  // the draw as a 64-bit word with its 52 bits on top (what Random.int multiplies)
  inline uint64_t lane() { return bits52() << 12; }
  // Random.float 1.0
  inline double uniform() {
    uint64_t bits = (0x3FFull << 52) | bits52();
    double d; std::memcpy(&d, &bits, 8);
    return d - 1.0;
  }
  // 1 + Random.float 1.0, in [1, 2): the raw mantissa pattern, no subtraction
  inline double uniform12() {
    uint64_t bits = (0x3FFull << 52) | bits52();
    double d; std::memcpy(&d, &bits, 8);
    return d;
  }
  // Random.int n
  inline uint64_t below(uint64_t n) {
    return (uint64_t)(((unsigned __int128)lane() * n) >> 64);
  }
};

}  // namespace og
