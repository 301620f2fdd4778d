#define CNF_TU_PREC CNF_PREC_FP16
#define CNF_TU_NAME tc2_forward_fp16
#include "tc2_fwd.inl"
CNF_DEFINE_SET_TRACE(set_trace_tc2_fwd_fp16)
