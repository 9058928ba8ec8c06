// conv_tc.cu -- tcgen05 (bf16) decoder convolutions.  Placeholder until the kernels land.
#include "common.cuh"
namespace iod {
int tc_supported(const Plan* p) { (void)p; set_error("IODINE_BF16: tensor-core path not built"); return 0; }
int tc_alloc(Plan*) { return 0; }
void tc_free(Plan*) {}
int tc_on_workspace(Plan*) { return 0; }
int tc_setup_weights(Plan*, const IodineWeights*, cudaStream_t) { return 0; }
int tc_launch_conv(Plan*, int, bool, const void*, const void*, void*, float*, cudaStream_t) { set_error("tc path not built"); return 1; }
int tc_launch_out4(Plan*, const void*, float*, cudaStream_t) { set_error("tc path not built"); return 1; }
int tc_launch_dgrad_in4(Plan*, const float*, const void*, void*, cudaStream_t) { set_error("tc path not built"); return 1; }
int tc_export_f32(Plan*, const void*, float*, size_t, cudaStream_t) { set_error("tc path not built"); return 1; }
}
