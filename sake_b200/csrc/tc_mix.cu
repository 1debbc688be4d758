// placeholder — replaced by the tcgen05 engine
#include "common.cuh"
namespace sake {
bool tc_supported(const Dims& d) { (void)d; return false; }
size_t tc_scratch_bytes(const Dims&, int, int, int) { return 0; }
int tc_mix_fwd(const Dims&, const SakeLayerParams&, const float*, const float*, const Saved&, void*, int, cudaStream_t) { set_error("tcgen05 engine not built"); return SAKE_EUNSUPPORTED; }
int tc_mix_bwd(const Dims&, const SakeLayerParams&, const float*, const float*, const Saved&, const BwdScratch&, float*, void*, int, cudaStream_t) { set_error("tcgen05 engine not built"); return SAKE_EUNSUPPORTED; }
int tc_selftest(float* e, cudaStream_t) { if (e) *e = -1.f; set_error("tcgen05 engine not built"); return SAKE_EUNSUPPORTED; }
}
