#include <string>
#include "../../include/ullsam.h"
#include "api_internal.h"
struct ullsam_handle { std::string err; };
using namespace ullsam;
extern "C" {
const char* ullsam_last_error(const ullsam_handle* h) { return h ? h->err.c_str() : g_thread_error.c_str(); }
int ullsam_create(const ullsam_config*, int, ullsam_handle**) { return op_fail(ULLSAM_ERR_STATE, "not implemented"); }
void ullsam_destroy(ullsam_handle*) {}
int ullsam_set_weight(ullsam_handle*, const char*, const void*, int, const int64_t*, int) { return op_fail(ULLSAM_ERR_STATE, "not implemented"); }
int ullsam_missing_weights(const ullsam_handle*, int) { return -1; }
size_t ullsam_workspace_bytes(const ullsam_handle*, int, int) { return 0; }
int ullsam_encoder_forward(ullsam_handle*, const void*, int, int, void*, int, void*) { return op_fail(ULLSAM_ERR_STATE, "not implemented"); }
int ullsam_decoder_forward(ullsam_handle*, const void*, int, const void*, const void*, int, int, const void*, int, int, int, void*, void*, void*) { return op_fail(ULLSAM_ERR_STATE, "not implemented"); }
int ullsam_decoder_predict_masks(ullsam_handle*, const void*, int, const void*, const void*, int, int, const void*, int, int, void*, void*, void*) { return op_fail(ULLSAM_ERR_STATE, "not implemented"); }
int ullsam_transformer_forward(ullsam_handle*, const void*, int, const void*, const void*, int, void*, void*, void*) { return op_fail(ULLSAM_ERR_STATE, "not implemented"); }
int ullsam_segment_host(ullsam_handle*, const float*, int, const float*, const float*, int, const float*, int, float*, float*, float*) { return op_fail(ULLSAM_ERR_STATE, "not implemented"); }
}
