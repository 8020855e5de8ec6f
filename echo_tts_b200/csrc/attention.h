// Flash-style attention over segmented keys [self | latent | text | speaker] (reference model.py:246-261), encoder
// self-attention (key mask or causal, model.py:141-154) and the DAC window-limited causal attention
// (autoencoder.py:698-702, 762-773): one tcgen05 kernel template over head dim 128 / 64 (attention_tc.cu).
#pragma once
#include <cuda_runtime.h>

#include "echo_b200.h"

namespace echo {
// Returns cudaErrorInvalidValue for descriptors the kernel cannot run (head dim other than 128 / 64, misaligned
// pointers / strides, or more than 7680 visible keys per 128-query tile) -- there is no second attention path.
cudaError_t attention_launch(const echo_attn_desc& d, cudaStream_t s);
}
