// thread-local last-error plumbing shared by host (.cpp) and device (.cu) translation units
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
int dcnet_set_error(int code, const char* fmt, ...);
void dcnet_count_launch(int n);
#ifdef __cplusplus
}
#endif
