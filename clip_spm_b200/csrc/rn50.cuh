// CLIP ModifiedResNet-50 frame encoder (models/clip_fsar.py:502-608, 396-500) -- interface used by model.cu.
#pragma once
#include <cuda_runtime.h>
#include <functional>
#include <string>

namespace spm {

struct Rn50;
// looks a reference state_dict tensor up by name, checks its element count, returns the device pointer
using WeightGetter = std::function<int(const std::string&, long long, const float**)>;

int rn50_create(Rn50** out, cudaStream_t st, int sms, const WeightGetter& get);
// images [F,3,224,224] fp32 NCHW -> feats [F,1024] fp32
int rn50_encode(Rn50* r, cudaStream_t st, const float* images, int n_frames, float* feats_out);
void rn50_destroy(Rn50* r);

}  // namespace spm
