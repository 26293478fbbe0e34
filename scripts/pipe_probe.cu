#include <cstdio>
#include <cstdint>
typedef uint32_t u32; typedef uint64_t u64;
// MODE: 0 add.u32, 1 add.cc+addc (64-bit add as two ops), 2 mad.lo.u32, 3 mad.wide.u32, 4 iadd3 3-input (a+b+c),
// 5 mad.lo + add interleaved, 6 mad.wide + add.cc/addc interleaved, 7 setp+selp pair, 8 lop3 (xor), 9 shf
template<int MODE> __global__ void probe(u64* out, int iters, u32 a, u32 b){
  u32 x[16], y[16];
  #pragma unroll
  for(int j=0;j<16;j++){ x[j]=threadIdx.x*77+j+a; y[j]=threadIdx.x*13+j*b; }
  __syncthreads();
  long long t0=clock64();
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int j=0;j<16;j++){
      if(MODE==0) asm volatile("add.u32 %0, %0, %1;":"+r"(x[j]):"r"(y[j]));
      if(MODE==1) asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;":"+r"(x[j]),"+r"(y[j]):"r"(a),"r"(b));
      if(MODE==2) asm volatile("mad.lo.u32 %0, %0, %1, %2;":"+r"(x[j]):"r"(a),"r"(y[j]));
      if(MODE==3){ u64 t=((u64)y[j]<<32)|x[j]; asm volatile("mad.wide.u32 %0, %1, %2, %0;":"+l"(t):"r"(a),"r"(b)); x[j]=(u32)t; y[j]=(u32)(t>>32);}
      if(MODE==4) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}":"+r"(x[j]):"r"(y[j]),"r"(a));
      if(MODE==5){ asm volatile("mad.lo.u32 %0, %0, %1, %2;":"+r"(x[j]):"r"(a),"r"(b)); asm volatile("add.u32 %0, %0, %1;":"+r"(y[j]):"r"(a)); }
      if(MODE==6){ u64 t=((u64)y[j]<<32)|x[j]; asm volatile("mad.wide.u32 %0, %1, %2, %0;":"+l"(t):"r"(a),"r"(b)); x[j]=(u32)t; y[j]=(u32)(t>>32); }
      if(MODE==10){ asm volatile("mad.lo.u32 %0, %0, %1, %2;":"+r"(x[j]):"r"(a),"r"(b)); }
      if(MODE==10 || MODE==11 || MODE==12){ u32 p=y[j], q=y[(j+1)&15]; asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;":"+r"(p),"+r"(q):"r"(a),"r"(b)); y[j]=p; y[(j+1)&15]=q; }
      if(MODE==11){ u64 t; asm volatile("mul.wide.u32 %0, %1, %2;":"=l"(t):"r"(x[j]),"r"(a)); x[j]=(u32)t^(u32)(t>>32); }
      if(MODE==13){ u64 t; asm volatile("mul.wide.u32 %0, %1, %2;":"=l"(t):"r"(x[j]),"r"(a)); x[j]=(u32)t^(u32)(t>>32); }
      if(MODE==14){ asm volatile("{.reg .pred p; setp.ge.u32 p, %0, %1; selp.u32 %0, %2, %0, p;}":"+r"(x[j]):"r"(a),"r"(b)); asm volatile("mad.lo.u32 %0, %0, %1, %2;":"+r"(y[j]):"r"(a),"r"(b)); }
      if(MODE==15){ asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}":"+r"(x[j]):"r"(y[j]),"r"(a)); asm volatile("mad.lo.u32 %0, %0, %1, %2;":"+r"(y[j]):"r"(a),"r"(b)); }
      if(MODE==20){ u64 t; asm volatile("mul.wide.u32 %0, %1, %2;":"=l"(t):"r"(x[j]),"r"(a)); x[j]=(u32)(t>>32); y[j]+=(u32)t; }
      if(MODE==21){ asm volatile("mul.hi.u32 %0, %0, %1;":"+r"(x[j]):"r"(a)); }
      if(MODE==22){ asm volatile("mul.lo.u32 %0, %0, %1;":"+r"(x[j]):"r"(a)); }
      if(MODE==23){ u64 t=((u64)y[j]<<32)|x[j]; asm volatile("mad.wide.u32 %0, %1, %2, %0;":"+l"(t):"r"(x[j]),"r"(a)); x[j]=(u32)t; y[j]=(u32)(t>>32); }
      if(MODE==24){ asm volatile("mad.hi.u32 %0, %0, %1, %2;":"+r"(x[j]):"r"(a),"r"(y[j])); }
      if(MODE==30){ u64 t=((u64)y[j]<<32)|x[j]; double d; asm volatile("cvt.rn.f64.u64 %0, %1;":"=d"(d):"l"(t)); u64 r; asm volatile("mov.b64 %0, %1;":"=l"(r):"d"(d)); x[j]=(u32)r; y[j]^=(u32)(r>>32); }
      if(MODE==31){ u64 t=((u64)(y[j]&0x3ff00000u|0x40000000u)<<32)|x[j]; double d; asm volatile("mov.b64 %0, %1;":"=d"(d):"l"(t)); u64 r; asm volatile("cvt.rzi.u64.f64 %0, %1;":"=l"(r):"d"(d)); x[j]=(u32)r; y[j]+=(u32)(r>>32); }
      if(MODE==32){ u64 t=((u64)(y[j]&0x3ff00000u|0x40000000u)<<32)|x[j]; double d; asm volatile("mov.b64 %0, %1;":"=d"(d):"l"(t)); asm volatile("mul.f64 %0, %0, %0;":"+d"(d)); u64 r; asm volatile("mov.b64 %0, %1;":"=l"(r):"d"(d)); x[j]=(u32)r; y[j]=(u32)(r>>32); }
      if(MODE==33){ float f; asm volatile("cvt.rn.f32.u32 %0, %1;":"=f"(f):"r"(x[j])); asm volatile("mov.b32 %0, %1;":"=r"(x[j]):"f"(f)); }
      if(MODE==7) asm volatile("{.reg .pred p; setp.ge.u32 p, %0, %1; selp.u32 %0, %2, %0, p;}":"+r"(x[j]):"r"(y[j]),"r"(a));
      if(MODE==8) asm volatile("xor.b32 %0, %0, %1;":"+r"(x[j]):"r"(y[j]));
      if(MODE==9) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;":"+r"(x[j]):"r"(y[j]));
    }
  }
  long long t1=clock64();
  u32 r=0;
  #pragma unroll
  for(int j=0;j<16;j++) r^=x[j]^y[j];
  if(r==0x12345u) out[1]=r;
  if(threadIdx.x==0 && blockIdx.x==0) out[0]=(u64)(t1-t0);
}
template<int MODE> void run(const char* name, int ops_per){
  u64* d; cudaMalloc(&d,16); int iters=4096;
  probe<MODE><<<148,1024>>>(d,iters,3,5); cudaDeviceSynchronize();
  probe<MODE><<<148,1024>>>(d,iters,3,5); cudaDeviceSynchronize();
  u64 h[2]; cudaMemcpy(h,d,16,cudaMemcpyDeviceToHost);
  double per_clk_sm = 1024.0*iters*16*ops_per/(double)h[0];
  printf("%-28s %8.1f thread-ops/clk/SM  (%.3f warp-inst/clk/SMSP)  err=%s\n",name,per_clk_sm,per_clk_sm/128.0,cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}
int main(){
  run<0>("add.u32",1); run<1>("add.cc+addc (2 ops)",2); run<2>("mad.lo.u32",1); run<3>("mad.wide.u32",1);
  run<4>("add3 (2 adds->IADD3)",1); run<5>("mad.lo + add (2 ops)",2); run<20>("mul.wide (+1 add) ",1); run<21>("mul.hi.u32",1); run<22>("mul.lo.u32",1); run<23>("mad.wide acc64 (var)",1); run<24>("mad.hi.u32",1); run<30>("cvt f64<-u64",1); run<31>("cvt u64<-f64",1); run<32>("mul.f64",1); run<33>("cvt f32<-u32",1); run<7>("setp+selp (2 ops)",2); run<12>("add.cc+addc alone (2 ops)",2); run<10>("mad.lo + add.cc+addc (3 ops)",3); run<13>("mul.wide+xor (2 ops)",2); run<11>("mul.wide+xor + add.cc+addc (4)",4); run<14>("setp+selp + mad.lo (3 ops)",3); run<15>("iadd3(3in) + mad.lo (2 ops)",2); run<8>("xor",1); run<9>("shf",1);
}
