// seqik_common.h -- error plumbing shared by the translation units of libseqik_sm100.so
#pragma once
int seqik_fail(int code, const char* fmt, const char* a = "");
int seqik_check_launch(const char* what);
