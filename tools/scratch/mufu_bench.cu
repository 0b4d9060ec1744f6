// Microbenchmark: achievable exp2 throughput per SM for softmax-like instruction mixes.
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
__device__ __forceinline__ float ex2(float x){float y; asm volatile("ex2.approx.ftz.f32 %0, %1;":"=f"(y):"f"(x)); return y;}
template<int MODE>
__global__ void k(float* out, const float* in, int iters, float c, float nm, long long* cyc){
    float s[64];
    #pragma unroll
    for(int i=0;i<64;++i) s[i]=in[(threadIdx.x*64+i)&1023];
    float r0=0,r1=0; uint32_t acc=0;
    __syncthreads();
    long long t0=clock64();
    for(int it=0;it<iters;++it){
        #pragma unroll
        for(int i=0;i<32;++i){
            if(MODE==0){ r0+=ex2(s[2*i]); r1+=ex2(s[2*i+1]); }
            else {
                float p0=ex2(fmaf(s[2*i],c,nm)), p1=ex2(fmaf(s[2*i+1],c,nm));
                r0+=p0; r1+=p1;
                __nv_bfloat162 v=__floats2bfloat162_rn(p0,p1); acc^=*reinterpret_cast<uint32_t*>(&v);
            }
        }
        nm+=1e-7f;
    }
    long long t1=clock64();
    out[blockIdx.x*blockDim.x+threadIdx.x]=r0+r1+__uint_as_float(acc);
    if(threadIdx.x==0) cyc[blockIdx.x]=t1-t0;
}
int main(){
    float *in,*out; long long* cyc; cudaMalloc(&in,4096); cudaMalloc(&out,1<<22); cudaMalloc(&cyc,148*8*8);
    cudaMemset(in,0,4096);
    for(int mode=0;mode<2;++mode) for(int warps: {4,8,16,32}){
        int iters=2000;
        if(mode==0) k<0><<<148,warps*32>>>(out,in,iters,1.f,0.f,cyc); else k<1><<<148,warps*32>>>(out,in,iters,1.f,0.f,cyc);
        cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h,cyc,sizeof(h),cudaMemcpyDeviceToHost);
        double c=0; for(int i=0;i<148;++i) c+=h[i]; c/=148;
        double exps=(double)warps*32*64*iters;
        printf("mode %d warps/SM %2d: %.0f cycles, %.2f exp/clk/SM\n",mode,warps,c,exps/c);
    }
    return 0;
}
