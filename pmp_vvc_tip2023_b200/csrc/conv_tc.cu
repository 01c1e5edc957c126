// placeholder until the tcgen05 engine lands (same translation unit name)
#include "handle.cuh"
#include "kernels.cuh"
namespace pmp {
bool tc_supported(int, int, int, int, int) { return false; }
size_t tc_packed_elems(int, int, int) { return 0; }
void pack_tc_weights(const float *, int, int, int, int, int, bool, uint16_t *) {}
int conv_tc(Handle *, const TcConvArgs &, int, cudaStream_t) { set_error("TC engine not built"); return PMP_ERR_UNSUPPORTED; }
int pool2_split(Handle *, const Act &, const Act &, const Act &, int, cudaStream_t) { set_error("TC engine not built"); return PMP_ERR_UNSUPPORTED; }
}
extern "C" int pmp_selftest_conv(pmp_handle *, int, int, int, int, int, int, double *, double *, double *, double *) { return PMP_ERR_UNSUPPORTED; }
