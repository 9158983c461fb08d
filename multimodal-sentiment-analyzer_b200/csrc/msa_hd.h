// Function qualifiers shared by the device build (nvcc) and the CPU emulation build (g++, tests/emu).
#pragma once
#ifdef __CUDACC__
#define MSA_FN __device__ __forceinline__
#define MSA_KFN __device__
#else
#define MSA_FN inline
#define MSA_KFN
#endif
