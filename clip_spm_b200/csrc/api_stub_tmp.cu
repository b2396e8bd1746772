// TEMPORARY: stubs for entry points not implemented yet (removed as the real ones land)
#include "../../include/clipspm_b200.h"
#include "api_common.cuh"
#define NI spm::set_error("not implemented yet"); return 1;
extern "C" {
int spm_create(const spm_config*, spm_handle**) { NI }
int spm_destroy(spm_handle*) { NI }
int spm_load_weights(spm_handle*, void*, int, const char* const*, const void* const*, const int64_t*) { NI }
int spm_set_text_features(spm_handle*, void*, const float*, int, int) { NI }
int spm_encode_frames(spm_handle*, void*, const float*, int, float*) { NI }
int spm_head(spm_handle*, void*, int, int, int, const float*, const float*, const float*, const float*, const float*, float*, float*) { NI }
int spm_forward(spm_handle*, void*, int, int, int, const float*, const float*, const float*, const float*, const float*, float*, float*) { NI }
int spm_eval(spm_handle*, void*, int, int, int, const float*, const float*, const float*, const float*, const float*, const int64_t*, float, float*, float*, float*, float*, int32_t*) { NI }
int spm_eval_host(spm_handle*, int, int, int, const float*, const float*, const float*, const float*, const float*, const int64_t*, float, float*, float*, float*, float*, int32_t*) { NI }
int spm_otam_distance(void*, int, int, int, int, int, const float*, const float*, int, float, float, float*) { NI }
}
