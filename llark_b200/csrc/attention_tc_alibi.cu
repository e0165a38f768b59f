// ALiBi instantiation of the tcgen05 attention kernels for the CLAP + MPT-1B variant: the same
// source as attention_tc.cu compiled with LLARK_ATC_ALIBI, i.e. scores = q.k / sqrt(D) + slope_h * (key - (S - 1)) (the bias
// of m2t/llava/model/mpt/attention.py:468-484, so the LSE matches the reference's too) and no RoPE.
// STAGED: compiled for sm_100a, not yet run on a B200 (GPU tests gated by LLARK_STAGED=1).
#define LLARK_ATC_ALIBI 1
#define LLARK_ATC_NS atc_alibi
#define LLARK_ATC_FWD_LAUNCH attn_alibi_fwd_tc_launch
#define LLARK_ATC_BWD_LAUNCH attn_alibi_bwd_tc_launch
#include "attention_tc.cu"
