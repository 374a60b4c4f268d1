// Internal (non-exported) entry points shared between translation units.
#pragma once
#include "common.cuh"

int mtus_gemm_simt(const mtus_gemm_desc* d, const EpiParams& ep, cudaStream_t st);
int mtus_gemm_tc2(const mtus_gemm_desc* d, cudaStream_t st);
bool mtus_gemm_tc2_supported(const mtus_gemm_desc* d);
