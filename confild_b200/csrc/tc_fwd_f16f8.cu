#define CNF_TU_PREC CNF_PREC_F16F8
#define CNF_TU_NAME tc_forward_f16f8
#include "tc_fwd.inl"
CNF_DEFINE_SET_TRACE(set_trace_tc_fwd_f16f8)
