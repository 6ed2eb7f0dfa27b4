// alac_kernels_depth.cu -- instantiates every encode and decode kernel for ONE bit depth (-DALAC_INSTANTIATE_DEPTH=16|20|24|32).
// The Makefile compiles it four times, in parallel; alac_engine.cu only sees `extern template` declarations of the launchers.
#ifndef ALAC_INSTANTIATE_DEPTH
#error "compile with -DALAC_INSTANTIATE_DEPTH=16|20|24|32"
#endif
#include "alac_decode.cuh"
#include "alac_encode.cuh"

namespace alacb {
template uint32_t enc_launch_search<ALAC_INSTANTIATE_DEPTH>(cudaStream_t, const EncArgs &, uint32_t, uint32_t, bool, bool, const JobLists *, cudaEvent_t *, int);
template void enc_launch_assemble<ALAC_INSTANTIATE_DEPTH>(cudaStream_t, const AsmArgs &);
template void dec_configure<ALAC_INSTANTIATE_DEPTH>();
template uint32_t dec_launch_main<ALAC_INSTANTIATE_DEPTH>(cudaStream_t, const DecArgs &, cudaEvent_t *);
}  // namespace alacb
