// Minimal stand-in for the OptiX 7 SDK header so that OWL's device/common headers parse.
#pragma once
#include <cuda_runtime.h>
typedef unsigned long long OptixTraversableHandle;
typedef unsigned long long CUdeviceptr_stub;
#ifdef __CUDACC__
static __forceinline__ __device__ uint3 optixGetLaunchIndex() { return make_uint3(0,0,0); }
static __forceinline__ __device__ uint3 optixGetLaunchDimensions() { return make_uint3(1,1,1); }
static __forceinline__ __device__ unsigned long long optixGetSbtDataPointer() { return 0ull; }
static __forceinline__ __device__ unsigned int optixGetPayload_0() { return 0u; }
static __forceinline__ __device__ unsigned int optixGetPayload_1() { return 0u; }
#endif
#ifdef __CUDACC__
static __forceinline__ __device__ float saturate(const float f) { return __saturatef(f); }
#endif
