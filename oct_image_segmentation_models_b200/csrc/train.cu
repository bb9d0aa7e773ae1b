// Training half of the C ABI (forward with batch-stat BN + dropout, weighted CE, backward,
// NCCL gradient all-reduce, Keras-Adam).  Filled in after the inference path is parity-green.
#include "net.cuh"

using namespace octseg;

extern "C" {

void octseg_train_free(octseg_net *net) { (void)net; }

int32_t octseg_train_begin(octseg_net *, const octseg_train_config *, const float *) {
  set_error("training path not built yet");
  return 1;
}
int32_t octseg_comm_unique_id(uint8_t *) { set_error("training path not built yet"); return 1; }
int32_t octseg_comm_init(octseg_net *, const uint8_t *, int32_t, int32_t) {
  set_error("training path not built yet");
  return 1;
}
int32_t octseg_train_step_host(octseg_net *, const void *, int32_t, const uint8_t *, int32_t, int32_t,
                               int32_t, const uint8_t *, float *) {
  set_error("training path not built yet");
  return 1;
}
int32_t octseg_train_step_device(octseg_net *, const void *, int32_t, const uint8_t *, int32_t, int32_t,
                                 int32_t, const uint8_t *, float *, void *) {
  set_error("training path not built yet");
  return 1;
}
int32_t octseg_get_grad(octseg_net *, int32_t, float *, int64_t) {
  set_error("training path not built yet");
  return 1;
}
}
