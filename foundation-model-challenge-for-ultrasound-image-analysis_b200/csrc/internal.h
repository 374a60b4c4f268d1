// Internal (non-exported) entry points shared between translation units.
#pragma once
#include "common.cuh"

int mtus_gemm_simt(const mtus_gemm_desc* d, const EpiParams& ep, cudaStream_t st);
int mtus_gemm_tc2(const mtus_gemm_desc* d, cudaStream_t st);
bool mtus_gemm_tc2_supported(const mtus_gemm_desc* d);

// tcgen05 window attention forward (attention_tc.cu): windows <= 64 tokens, bf16, unpadded maps, even head count
bool mtus_window_attn_tc_eligible(int B, int H, int W, int C, int heads, int wh, int ww, int sh, int sw, int dtype);
int mtus_window_attn_tc_fwd(const void* qkv, const float* rel_table, void* out, float* lse, int B, int H, int W, int C, int heads,
                            int wh, int ww, int sh, int sw, cudaStream_t st);
