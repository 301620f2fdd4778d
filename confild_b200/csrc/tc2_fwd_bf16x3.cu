#define CNF_TU_PREC CNF_PREC_BF16X3
#define CNF_TU_NAME tc2_forward_bf16x3
#include "tc2_fwd.inl"
CNF_DEFINE_SET_TRACE(set_trace_tc2_fwd_bf16x3)
