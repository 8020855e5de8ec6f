// Flash-style attention over segmented keys [self | latent | text | speaker] (reference model.py:246-261).
#pragma once
#include <cuda_runtime.h>

#include "echo_b200.h"

namespace echo {
cudaError_t attention_launch(const echo_attn_desc& d, cudaStream_t s);
}
