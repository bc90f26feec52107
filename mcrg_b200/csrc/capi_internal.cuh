// What the other translation units of the C ABI need from capi.cu: error reporting and a few context fields.
#pragma once
#include <cuda_runtime.h>

#include "../../include/mcrg_b200.h"
#include "kernels.cuh"

int mcrg_fail(int code, const char *fmt, ...);
int mcrg_ctx_device(const mcrg_ctx *c);
cudaStream_t mcrg_ctx_stream(const mcrg_ctx *c);
void *mcrg_ctx_comm(const mcrg_ctx *c);   // ncclComm_t or nullptr
void *mcrg_ctx_limbs(const mcrg_ctx *c);  // device buffer of 4*N_SLOTS int64 owned by the communicator set-up
void mcrg_ctx_set_comm(mcrg_ctx *c, void *comm, void *limbs);
