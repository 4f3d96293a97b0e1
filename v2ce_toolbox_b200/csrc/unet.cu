// TEMPORARY stubs (replaced by the real stage-1 implementation in the next commit).
#include "common.cuh"
using namespace v2ce;
#define STUB return set_error(V2CE_ERR_STATE, "stage 1 not built yet")
extern "C" int v2ce_model_create(v2ce_model**, int) { STUB; }
extern "C" int v2ce_model_destroy(v2ce_model*) { STUB; }
extern "C" int v2ce_model_set_tensor(v2ce_model*, const char*, const float*, const int64_t*, int32_t) { STUB; }
extern "C" int v2ce_model_finalize(v2ce_model*) { STUB; }
extern "C" int v2ce_model_workspace_bytes(const v2ce_model*, int32_t, int32_t, int32_t, int32_t, size_t*) { STUB; }
extern "C" int v2ce_model_forward(v2ce_model*, const float*, float*, int32_t, int32_t, int32_t, int32_t, void*, size_t, void*) { STUB; }
extern "C" int v2ce_model_last_sigmas(const v2ce_model*, float*) { STUB; }
extern "C" int v2ce_model_call_count(const v2ce_model*, int64_t*) { STUB; }
extern "C" int v2ce_model_sn_advance(v2ce_model*, int32_t, void*) { STUB; }
extern "C" int v2ce_model_last_launches(const v2ce_model*, int32_t*) { STUB; }
extern "C" int v2ce_conv3d_bf16(const void*, int32_t, int32_t, int32_t, const void*, int32_t, int32_t, int32_t, int32_t, int32_t, int32_t, int32_t, const float*, int32_t, const float*, const float*, const void*, int32_t, void*, void*) { STUB; }
