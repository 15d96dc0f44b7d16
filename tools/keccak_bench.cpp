// Dev microbenchmark: scalar vs AVX-512 single-state vs AVX-512 x8 Keccak-f[1600] on the host CPU.
//   g++ -O3 -march=x86-64-v3 -std=c++17 -I bulletproof_gadgets_b200/csrc tools/keccak_bench.cpp bulletproof_gadgets_b200/csrc/keccak_x8_native.cpp -x c++ bulletproof_gadgets_b200/csrc/merlin.cpp -o /tmp/keccak_bench -lpthread
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <time.h>
namespace bpg { void keccak_f1600_reference(uint64_t st[25]); void keccak_f1600(uint64_t st[25]); void keccak_f1600_avx512(uint64_t st[25]); void keccak_f1600_x8(uint64_t st[25][8]); }
static double now(){ timespec ts; clock_gettime(CLOCK_MONOTONIC,&ts); return ts.tv_sec+ts.tv_nsec*1e-9; }
int main(){
  uint64_t a[25], b[25]; int bad=0;
  for (int t=0;t<1000;t++){ for(int i=0;i<25;i++) a[i]=b[i]=(uint64_t)(t+1)*0x9e3779b97f4a7c15ULL*(i+3) ^ (a[i]>>7); bpg::keccak_f1600_reference(a); bpg::keccak_f1600_avx512(b); if(memcmp(a,b,200)) bad++; }
  printf("mismatches %d\n", bad);
  int N=2000000; static uint64_t v[25][8];
  for(int rep=0;rep<3;rep++){ double t0=now(); for(int i=0;i<N;i++) bpg::keccak_f1600(a); double t1=now(); for(int i=0;i<N;i++) bpg::keccak_f1600_avx512(b); double t2=now(); for(int i=0;i<N;i++) bpg::keccak_f1600_x8(v); double t3=now(); printf("x8 %.1f ns per vector permutation; ", (t3-t2)/N*1e9);
  printf("dispatch(keccak_f1600) %.1f ns  avx512-1x %.1f ns  (%llx %llx)\n", (t1-t0)/N*1e9, (t2-t1)/N*1e9, (unsigned long long)a[0], (unsigned long long)b[0]); }
}
