// One thread per cycle of the sorted order (see witgen.cu): the kernel that instantiates the generated rv32im step
// function `step_Top` (gen/witgen_rv32im.inc, tools/gen_witgen.py). Own translation unit because ptxas needs minutes for it.
#include <cuda_runtime.h>

#include "witgen_rt.cuh"

namespace r0wg {
namespace {   // internal linkage: both step translation units instantiate the same generated functions
#include "gen/witgen_rv32im.inc"
}  // namespace

#ifndef WG_MIN_BLOCKS
#define WG_MIN_BLOCKS 6   // measured at po2 = 20 (tools/witgen_variants.sh): 1 -> 4.30 ms, 6 -> 2.54, 8 -> 2.31, 10 -> 2.21; but under the 64- / 48-register caps of 8 / 10 the -O1 build mis-computes the all-instruction guest (tests/test_gpu_witgen.py [all_insn], eqz failure) while the host build of the same text is clean under ASan + UBSan: 6 is the tightest setting the whole suite passes with
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
