// placeholder until the tcgen05 chain lands (replaced below in this round)
#include "cb_common.cuh"
struct cb_tc_chain { int unused; };
extern "C" int32_t cb_tc_chain_create(const cb_mlp_desc*, int32_t, cb_tc_chain**) {
  cb::set_error("tensor-core chain not built yet");
  return CB_ERR_ARG;
}
extern "C" void cb_tc_chain_destroy(cb_tc_chain*) {}
extern "C" int32_t cb_tc_chain_set_weights(cb_tc_chain*, const float* const*, const float* const*, void*) {
  cb::set_error("tensor-core chain not built yet");
  return CB_ERR_ARG;
}
extern "C" int32_t cb_tc_chain_forward(cb_tc_chain*, const float*, int64_t, float*, void*) {
  cb::set_error("tensor-core chain not built yet");
  return CB_ERR_ARG;
}
