#include "conv_umma.cuh"
namespace cdan {
struct UmmaPack { int dummy; };
int umma_pack_create(const float*, const float*, int, int, int, int, UmmaPack** out) { *out = nullptr; return fail("tcgen05 conv not built yet"); }
void umma_pack_destroy(UmmaPack*) {}
bool conv_umma_supported(const ConvDesc&) { return false; }
int conv_umma_launch(const ConvDesc&, const UmmaPack&, cudaStream_t) { return fail("tcgen05 conv not built yet"); }
}
