// Flash-style attention over segmented keys [self | latent | text | speaker] (reference model.py:246-261).
#pragma once
#include <cuda_runtime.h>

#include "echo_b200.h"

namespace echo {
// Dispatch: head_dim 128 without causal segments -> tcgen05 kernel (attention_tc.cu); causal / window / head_dim 64
// (encoders, DAC post_module; < 1 % of a request) -> mma.sync kernel (attention.cu). ECHO_ATTN_LEGACY=1 forces the latter.
cudaError_t attention_launch(const echo_attn_desc& d, cudaStream_t s);
bool attention_tc_supported(const echo_attn_desc& d);
cudaError_t attention_tc_launch(const echo_attn_desc& d, cudaStream_t s);
}
