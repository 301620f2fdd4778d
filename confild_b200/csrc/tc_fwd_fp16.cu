#define CNF_TU_PREC CNF_PREC_FP16
#define CNF_TU_NAME tc_forward_fp16
#include "tc_fwd.inl"
CNF_DEFINE_SET_TRACE(set_trace_tc_fwd_fp16)
