// placeholder until the tcgen05 kernel lands: reports "not available" so the engine uses the SIMT conv
#include "conv_tc.h"
struct TcWeights { int n; };
bool tc_available() { return false; }
cudaError_t tc_alloc_weights(TcWeights** w, int nlayers) { *w = new TcWeights{nlayers}; return cudaSuccess; }
void tc_free_weights(TcWeights* w) { delete w; }
cudaError_t tc_prepare_weights(TcWeights*, int, const float*, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t tc_conv64(TcWeights*, const TcConvArgs&, cudaStream_t) { return cudaErrorNotSupported; }
