// SASS probe: which packed/DPX/video intrinsics are single hardware instructions on sm_100a?
#include <cuda_fp16.h>
#include <cstdint>
extern "C" {
__global__ void k_viaddmax_s32(int* o, int a, int b, int c){ o[threadIdx.x] = __viaddmax_s32(a+threadIdx.x,b,c); }
__global__ void k_viaddmax_s32_relu(int* o, int a, int b, int c){ o[threadIdx.x] = __viaddmax_s32_relu(a+threadIdx.x,b,c); }
__global__ void k_vimax3_s32(int* o, int a, int b, int c){ o[threadIdx.x] = __vimax3_s32(a+threadIdx.x,b,c); }
__global__ void k_vimax3_s32_relu(int* o, int a, int b, int c){ o[threadIdx.x] = __vimax3_s32_relu(a+threadIdx.x,b,c); }
__global__ void k_vimin_s32_relu(int* o, int a, int b){ o[threadIdx.x] = __vimin_s32_relu(a+threadIdx.x,b); }
__global__ void k_viaddmax_s16x2(unsigned* o, unsigned a, unsigned b, unsigned c){ o[threadIdx.x] = __viaddmax_s16x2(a+threadIdx.x,b,c); }
__global__ void k_viaddmax_s16x2_relu(unsigned* o, unsigned a, unsigned b, unsigned c){ o[threadIdx.x] = __viaddmax_s16x2_relu(a+threadIdx.x,b,c); }
__global__ void k_viaddmax_u16x2(unsigned* o, unsigned a, unsigned b, unsigned c){ o[threadIdx.x] = __viaddmax_u16x2(a+threadIdx.x,b,c); }
__global__ void k_vimax3_s16x2(unsigned* o, unsigned a, unsigned b, unsigned c){ o[threadIdx.x] = __vimax3_s16x2(a+threadIdx.x,b,c); }
__global__ void k_vimax3_s16x2_relu(unsigned* o, unsigned a, unsigned b, unsigned c){ o[threadIdx.x] = __vimax3_s16x2_relu(a+threadIdx.x,b,c); }
__global__ void k_vimin3_s16x2(unsigned* o, unsigned a, unsigned b, unsigned c){ o[threadIdx.x] = __vimin3_s16x2(a+threadIdx.x,b,c); }
__global__ void k_vimax_s16x2_relu(unsigned* o, unsigned a, unsigned b){ o[threadIdx.x] = __vimax_s16x2_relu(a+threadIdx.x,b); }
__global__ void k_vimin_s16x2_relu(unsigned* o, unsigned a, unsigned b){ o[threadIdx.x] = __vimin_s16x2_relu(a+threadIdx.x,b); }
__global__ void k_vibmax_s16x2(unsigned* o, unsigned a, unsigned b){ bool ph, pl; unsigned r = __vibmax_s16x2(a+threadIdx.x,b,&ph,&pl); o[threadIdx.x] = r + (ph?1:0) + (pl?2:0); }
__global__ void k_vmaxs2(unsigned* o, unsigned a, unsigned b){ o[threadIdx.x] = __vmaxs2(a+threadIdx.x,b); }
__global__ void k_vmaxu2(unsigned* o, unsigned a, unsigned b){ o[threadIdx.x] = __vmaxu2(a+threadIdx.x,b); }
__global__ void k_vadd2(unsigned* o, unsigned a, unsigned b){ o[threadIdx.x] = __vadd2(a+threadIdx.x,b); }
__global__ void k_vsub2(unsigned* o, unsigned a, unsigned b){ o[threadIdx.x] = __vsub2(a+threadIdx.x,b); }
__global__ void k_vaddus2(unsigned* o, unsigned a, unsigned b){ o[threadIdx.x] = __vaddus2(a+threadIdx.x,b); }
__global__ void k_vsubus2(unsigned* o, unsigned a, unsigned b){ o[threadIdx.x] = __vsubus2(a+threadIdx.x,b); }
__global__ void k_vcmpeq2(unsigned* o, unsigned a, unsigned b){ o[threadIdx.x] = __vcmpeq2(a+threadIdx.x,b); }
__global__ void k_vseteq2(unsigned* o, unsigned a, unsigned b){ o[threadIdx.x] = __vseteq2(a+threadIdx.x,b); }
__global__ void k_vmaxu4(unsigned* o, unsigned a, unsigned b){ o[threadIdx.x] = __vmaxu4(a+threadIdx.x,b); }
__global__ void k_vaddus4(unsigned* o, unsigned a, unsigned b){ o[threadIdx.x] = __vaddus4(a+threadIdx.x,b); }
__global__ void k_vsubus4(unsigned* o, unsigned a, unsigned b){ o[threadIdx.x] = __vsubus4(a+threadIdx.x,b); }
__global__ void k_vcmpeq4(unsigned* o, unsigned a, unsigned b){ o[threadIdx.x] = __vcmpeq4(a+threadIdx.x,b); }
__global__ void k_vadd4(unsigned* o, unsigned a, unsigned b){ o[threadIdx.x] = __vadd4(a+threadIdx.x,b); }
__global__ void k_hadd2(__half2* o, __half2 a, __half2 b){ o[threadIdx.x] = __hadd2(o[threadIdx.x],b); }
__global__ void k_hmax2(__half2* o, __half2 a, __half2 b){ o[threadIdx.x] = __hmax2(o[threadIdx.x],b); }
__global__ void k_hmin2(__half2* o, __half2 a, __half2 b){ o[threadIdx.x] = __hmin2(o[threadIdx.x],b); }
__global__ void k_hfma2_relu(__half2* o, __half2 a, __half2 b){ o[threadIdx.x] = __hfma2_relu(o[threadIdx.x],a,b); }
__global__ void k_heq2(__half2* o, __half2 a, __half2 b){ o[threadIdx.x] = __heq2(o[threadIdx.x],b); }
__global__ void k_fmax3(float* o, float a, float b){ o[threadIdx.x] = fmaxf(fmaxf(o[threadIdx.x],a),b); }
}
