// Launch accounting + optional per-kernel-class CUDA-event timing (bench.py's live roofline measurement).
// Disabled by default: a launch then costs one relaxed counter increment.  When a class is enabled with
// linr_prof_enable(mask), every launch of that class is bracketed by a cudaEvent pair on the launching stream.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace linr {

enum KClass {
    K_CONV88 = 0,   // conv27_kernel<8,8,0>  (forward ConvB/SConv inputs, and every 8->8 grad-input pass)
    K_CONV84,       // conv27_kernel<8,4,0>
    K_CONV48,       // conv27_kernel<4,8,0>
    K_CONV44,       // conv27_kernel<4,4,0>
    K_CONVBITS,     // conv27_kernel<8,8,1>  (ConvA of the LDFE blocks, occupancy-bit input)
    K_CONVHEAD,     // conv27_kernel<8,8,2>  (SConv_k + MLP_k + sigmoid + bits + CDF)
    K_BWDW88,       // conv27_bwd_w_kernel<8,8,0>
    K_BWDW84,
    K_BWDW44,
    K_BWDWBITS,     // conv27_bwd_w_kernel<8,8,1>
    K_PW,           // (unused since the pointwise convs run in conv epilogues; kept so class ids stay stable)
    K_PWBWDW,       // pointwise weight gradient
    K_HEADBWD,      // head_bwd_rows + head_bwd_w
    K_SCE,          // sce_fwd / sce_bwd / sce_finalize
    K_REDUCE,       // sum_groups, finalize_grad, bits_finalize
    K_ADAM,         // adam, param_quant, occ_set_stage
    K_COORD,        // coordinate-stage kernels of coords.cu (cub launches not counted)
    K_NCLASS
};

void prof_begin(int cls, int64_t units, cudaStream_t s);
void prof_end(int cls, cudaStream_t s);

struct ProfScope {
    int cls;
    cudaStream_t s;
    ProfScope(int c, int64_t units, cudaStream_t st) : cls(c), s(st) { prof_begin(c, units, st); }
    ~ProfScope() { prof_end(cls, s); }
};

}  // namespace linr
