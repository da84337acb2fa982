// Fused SMEM-resident pose-cell kernel (placeholder until the fused kernel lands: every plan
// uses the generic path).
#include "common.cuh"

int prs_pc_resident_supported(const prs_pc_plan*) { return 0; }

int prs_pc_resident_step(prs_pc_plan*, void*, const double*, int, const void*, long long*, void*, int*, cudaStream_t) {
  prs_set_error("resident path not available for this plan");
  return PRS_E_INVALID;
}
