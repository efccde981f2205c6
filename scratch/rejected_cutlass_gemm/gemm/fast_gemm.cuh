// fp32 GEMMs of the field MLP's Linear layers on the BF16 tensor cores (tcgen05) with the BF16x9 split: CUTLASS'
// sm_100 "FastF32" collective (cutlass/gemm/collective/sm100_mma_warpspecialized_emulated.hpp — TMA loads of the fp32
// tiles, a transform warp group that splits every operand into three bf16 terms, tcgen05.mma over the five product
// bands with fp32 accumulation in TMEM), instantiated here from the header tree this image vendors.  This is library
// code in the sense of the task (it counts like cuBLAS); what it buys over calling cuBLAS 12.9's own BF16x9 path is
// that cuBLAS brackets every GEMM with an inf/NaN scan of A and B (inf_patching::scan_AB_kernel, 26 us per call,
// 256 calls and 6.8 ms of a 36 ms training step) which a network whose activations are finite does not need.
#pragma once
#include "cutlass/cutlass.h"
#include "cute/tensor.hpp"
#include "cutlass/gemm/device/gemm_universal_adapter.h"
#include "cutlass/gemm/kernel/gemm_universal.hpp"
#include "cutlass/gemm/collective/collective_builder.hpp"
#include "cutlass/epilogue/collective/collective_builder.hpp"
#include "cutlass/util/packed_stride.hpp"

namespace nsvf_gemm {
using namespace cute;

template <class LayoutA, class LayoutB, class MmaTile, class Cluster, class KSched, class ESched>
struct FastGemm {
  using CollectiveEpilogue = typename cutlass::epilogue::collective::CollectiveBuilder<
      cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, MmaTile, Cluster,
      cutlass::epilogue::collective::EpilogueTileAuto, float, float,
      float, cutlass::layout::RowMajor, 4, float, cutlass::layout::RowMajor, 4, ESched>::CollectiveOp;
  using CollectiveMainloop = typename cutlass::gemm::collective::CollectiveBuilder<
      cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, float, LayoutA, 4, float, LayoutB, 4, float, MmaTile, Cluster,
      cutlass::gemm::collective::StageCountAutoCarveout<static_cast<int>(sizeof(typename CollectiveEpilogue::SharedStorage))>,
      KSched>::CollectiveOp;
  using GemmKernel = cutlass::gemm::kernel::GemmUniversal<Shape<int, int, int, int>, CollectiveMainloop, CollectiveEpilogue, void>;
  using Gemm = cutlass::gemm::device::GemmUniversalAdapter<GemmKernel>;
};

// D[l] (M x N, row-major, batch stride ldd_batch) = A[l] (M x K) * B[l] (K x N); strides in elements.
// Returns 0 on success, 10 + status otherwise.
template <class G>
int run(int M, int N, int K, int L, const float* A, long long a_batch, const float* B, long long b_batch, float* D,
        long long d_batch, void* ws, size_t ws_bytes, cudaStream_t stream) {
  using Gemm = typename G::Gemm;
  using StrideA = typename Gemm::GemmKernel::StrideA;
  using StrideB = typename Gemm::GemmKernel::StrideB;
  using StrideC = typename Gemm::GemmKernel::StrideC;
  using StrideD = typename Gemm::GemmKernel::StrideD;
  StrideA sa = cutlass::make_cute_packed_stride(StrideA{}, make_shape(M, K, L));
  StrideB sb = cutlass::make_cute_packed_stride(StrideB{}, make_shape(N, K, L));
  StrideC sc = cutlass::make_cute_packed_stride(StrideC{}, make_shape(M, N, L));
  StrideD sd = cutlass::make_cute_packed_stride(StrideD{}, make_shape(M, N, L));
  if (L > 1) {
    get<2>(sa) = a_batch;
    get<2>(sb) = b_batch;
    get<2>(sc) = d_batch;
    get<2>(sd) = d_batch;
  }
  typename Gemm::Arguments args{cutlass::gemm::GemmUniversalMode::kGemm, {M, N, K, L}, {A, sa, B, sb},
                                {{1.0f, 0.0f}, D, sc, D, sd}};
  Gemm gemm;
  cutlass::Status st = gemm.can_implement(args);
  if (st != cutlass::Status::kSuccess) return 10 + (int)st;
  if (Gemm::get_workspace_size(args) > ws_bytes) return 2;
  st = gemm.initialize(args, ws, stream);
  if (st != cutlass::Status::kSuccess) return 30 + (int)st;
  st = gemm.run(stream);
  if (st != cutlass::Status::kSuccess) return 50 + (int)st;
  return 0;
}

}  // namespace nsvf_gemm
