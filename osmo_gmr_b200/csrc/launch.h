// launch.h - internal launcher entry points shared between the kernel translation units and
// the C-ABI layer (api.cu).
#pragma once
#include <cuda_runtime.h>
#include "decode_unit.cuh"

namespace gmr1 {
cudaError_t launch_decode(int ch, const DecodeArgs &a, cudaStream_t st);
}
