// ullsam_b200 — helpers shared by the C-ABI translation units.
#pragma once

#include <string>

struct ullsam_handle;

namespace ullsam {

extern thread_local std::string g_thread_error;
int op_fail(int code, const char* msg);
// SM count of the current device; 0 if there is no CUDA device, -1 if it is not sm_100.
int device_sm_count();

}  // namespace ullsam
