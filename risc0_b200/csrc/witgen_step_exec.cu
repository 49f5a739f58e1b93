// One thread per cycle of the sorted order (see witgen.cu): the kernel that instantiates the generated rv32im step
// function `step_Top` (gen/witgen_rv32im.inc, tools/gen_witgen.py). Own translation unit because ptxas needs minutes for it.
#include <cuda_runtime.h>

#include "witgen_rt.cuh"

namespace r0wg {
namespace {   // internal linkage: both step translation units instantiate the same generated functions
#include "gen/witgen_rv32im.inc"
}  // namespace

#ifndef WG_MIN_BLOCKS
#define WG_MIN_BLOCKS 8   // measured at po2 = 20 (tools/witgen_const_check.sh, profiles/r2_witgen_out_params.log): 6 -> 2.66 ms, 8 -> 2.42, 10 -> 2.47.
// History: with aggregates returned BY VALUE from the non-inlined step functions, the 8 / 10 block builds (64 / 48
// registers) and the ptxas -O3 build mis-computed the all-instruction guest - bits of the first ToBits_16 result changed
// across the second call inside BitwiseAndU16. The generator now hands such results back through a reference
// (tools/gen_witgen.py returns_by_out_param); every one of those builds passes the whole witgen suite since.
#endif
__global__ void __launch_bounds__(128, WG_MIN_BLOCKS) k_step_exec(const WShared* s, const uint32_t* order, uint32_t begin, uint32_t count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint32_t cycle = order[begin + i];
  WCtx ctx{s, cycle, s->cycles[cycle].txn_idx};
  step_Top(ctx, BUF_DATA, BUF_GLOBAL);
}

void launch_step_exec(cudaStream_t stream, const WShared* s, const uint32_t* order, uint32_t begin, uint32_t count) {
  k_step_exec<<<(count + 127) / 128, 128, 0, stream>>>(s, order, begin, count);
}

}  // namespace r0wg
