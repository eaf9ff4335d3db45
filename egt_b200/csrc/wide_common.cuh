// wide_common.cuh -- compile-time geometry of the width-generic fused kernels (wide_fwd.cu, wide_bwd.cu).
#pragma once
#include "fused.h"      // encode_tmap_3d
#include "umma.cuh"

namespace egt {

// exponent budget of the un-normalised softmax, as in fused.h
constexpr float kWideSoftmaxBudget = 75.f;

// H heads of DK channels (model width D = H*DK), edge width DE, NG compute warpgroups, NS shared-memory stages.
template <int H_, int DK_, int DE_, int NG_, int NS_>
struct WideGeo {
  static constexpr int H = H_, DK = DK_, DE = DE_, NG = NG_, NS = NS_;
  static constexpr int D = H * DK;
  static constexpr int DKS = D / 16;                     // k-steps of Q K^T
  static constexpr int NQA = (D + 63) / 64;              // 64-channel swizzle atoms of a [128 x D] tile
  static constexpr int EGN = 2 * H;                      // columns of the [E|G] product of one key
  static constexpr int DEP = DE < 16 ? 16 : DE;          // columns of the e' / d x^ product of one key
  static constexpr int DEW = DE < 16 ? 16 : DE;          // K window of the e operand of one key
  static constexpr int KPB = 64 / DE;                    // keys per 128-byte TMA box row
  static constexpr int TK = NG > KPB ? NG : KPB;         // keys per tile (= per stage)
  static constexpr int KPG = TK / NG;                    // keys per group per tile
  static constexpr int NBOX = TK * DE / 64;              // [128 x 128 B] boxes per tile
  static constexpr int KV_ROWS = (TK * D * 2 + 127) & ~127;   // bytes of the K (or V) rows of a tile
  static constexpr int THREADS = 128 * (NG + 1);
  // setmaxnreg moves registers from the issuer / producer warpgroup to the compute warpgroups (only needed when
  // the launch leaves fewer than 112 per thread)
  static constexpr bool USE_SETMAXNREG = NG == 4;
  static constexpr int REG_COMPUTE = 112, REG_HELPER = 32;   // 512 x 112 + 128 x 32 == 640 x 96
  static_assert(H == 8 || H == 16, "heads");
  static_assert(D % 16 == 0 && D <= 128, "model width");
  static_assert(64 % DE == 0 && DE % 8 == 0, "edge width must divide a 128-byte row");
  static_assert(TK % NG == 0 && TK % KPB == 0, "tile geometry");
};

// forward: shared memory = Q | expanded K,V per group | stages | weights ...   (wide_fwd.cu recomputes the map)
template <class G>
struct WideFwdSmem {
  static constexpr int STAGE = (G::NBOX * 16384 + 2 * G::KV_ROWS + 1023) & ~1023;
  static constexpr int W = 2 * (2 * G::DEW * G::EGN * 2) + 2 * G::DEP * 32 + 1024;
  static constexpr int TOTAL = G::NQA * 16384 + G::NG * 3 * G::NQA * 2048 + G::NS * STAGE + W + 4096 + 256 + 256 + 4096 + 64 + 128 * G::H * 4;
};

struct WideFwdC5 : WideGeo<16, 8, 32, 4, 3> { static constexpr int FWD_SMEM = WideFwdSmem<WideGeo<16, 8, 32, 4, 3>>::TOTAL; };
struct WideFwdC1 : WideGeo<8, 8, 64, 2, 4> { static constexpr int FWD_SMEM = WideFwdSmem<WideGeo<8, 8, 64, 2, 4>>::TOTAL; };
struct WideFwdC3 : WideGeo<8, 12, 8, 4, 4> { static constexpr int FWD_SMEM = WideFwdSmem<WideGeo<8, 12, 8, 4, 4>>::TOTAL; };

}  // namespace egt
