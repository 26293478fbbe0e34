// ASAN/UBSAN fuzz of the host codec (run from the repo root):
//   g++ -g -O1 -fsanitize=address,undefined -std=c++17 -Ifhe_precompiles_b200/csrc -I/usr/local/cuda/include scripts/fuzz_codec.cpp \
//       fhe_precompiles_b200/csrc/codec.cpp fhe_precompiles_b200/csrc/context.cpp -o /tmp/fuzz -lz -ldl -lpthread -L/usr/local/cuda/lib64 -lcudart_static -lrt
//   ASAN_OPTIONS=detect_leaks=0 /tmp/fuzz
// Round 1: 3,600 mutated ciphertext / key blobs, 4,000 mutated + 1,700 header-bit-flipped structured frames and 2,000 random
// framings, no sanitizer report (codes 0 / 3 / 7 only).
#include "codec.h"
#include "kernels.h"
#include <cstdio>
#include <vector>
#include <random>
using namespace fheb;
namespace fheb { cudaError_t upload_constants(const DevConsts&, const DevTables&, const DevTwLow&){return cudaSuccess;} cudaError_t kernels_configure(){return cudaSuccess;} }
static std::vector<uint8_t> rd(const char* p){ FILE*f=fopen(p,"rb"); std::vector<uint8_t> b(1<<20); size_t n=fread(b.data(),1,b.size(),f); fclose(f); b.resize(n); return b; }
int main(){
  auto ct=rd("tests/golden/ct_i64_16_seed11.bin"); auto pk=rd("tests/data/public_key.bin"); auto sk=rd("tests/data/private_key.bin");
  std::mt19937_64 g(7); std::vector<uint64_t> w(kRkWords), w2(kRkWords); int hist[8]={0};
  auto mutate=[&](std::vector<uint8_t> b){ int k=g()%5; if(k==0) b.resize(g()%b.size()); else if(k==1){ for(int i=0;i<1+(int)(g()%8);i++) b[g()%b.size()]^=1<<(g()%8);} else if(k==2){ size_t off=g()%200; uint64_t v=g()>>(g()%64); for(int i=0;i<8&&off+i<b.size();i++) b[off+i]=(uint8_t)(v>>(8*i)); } else if(k==3){ for(int i=0;i<(int)(g()%64)+1;i++) b.push_back((uint8_t)g()); } else { size_t i=g()%(b.size()-16); for(int j=0;j<16;j++) b[i+j]=(uint8_t)g(); } return b; };
  for(int it=0;it<3000;it++){ auto m=mutate(ct); CipherView v; int rc=decode_ciphertext(Span{m.data(),m.size()},&v,w.data()); hist[rc&7]++; }
  // the same ciphertext re-written as a structured frame (codec.cpp zstd_pack40): the recogniser parses untrusted bytes too.
  // (agreement of the recogniser with libzstd on edited frames is tested in tests/test_formats.py)
  { CipherView v; if(decode_ciphertext(Span{ct.data(),ct.size()},&v,w.data())) return 1; std::vector<uint8_t> st; set_zstd_writer(1); encode_ciphertext(v,w.data(),&st);
    if(st.size()>=ct.size()) { printf("structured frame not smaller?\n"); return 1; }
    for(int it=0;it<4000;it++){ auto m=mutate(st); CipherView v2; int rc=decode_ciphertext(Span{m.data(),m.size()},&v2,w2.data()); hist[rc&7]++; }
    // header-region bit flips, exhaustively over the first 200 and last 16 bytes of the frame
    size_t frame=st.size()-82054; for(size_t pos=frame; pos<st.size(); pos++){ if(pos>frame+200 && pos+16<st.size()) continue; for(int bit=0;bit<8;bit++){ auto m=st; m[pos]^=1<<bit; CipherView v2; int rc=decode_ciphertext(Span{m.data(),m.size()},&v2,w2.data()); hist[rc&7]++; } } }
  for(int it=0;it<300;it++){ auto m=mutate(pk); bool has; int rc=decode_public_key(Span{m.data(),m.size()},w.data(),w2.data(),&has); hist[rc&7]++; auto m2=mutate(sk); rc=decode_private_key(Span{m2.data(),m2.size()},w.data()); hist[rc&7]++; }
  // framing
  for(int it=0;it<2000;it++){ std::vector<uint8_t> m(g()%64); for(auto&x:m) x=(uint8_t)g(); Span a,b,c; unpack_binary_operation(Span{m.data(),m.size()},&a,&b,&c); unpack_two_arguments(Span{m.data(),m.size()},&a,&b); uint16_t pl[kN]; for(int k=0;k<4;k++) encode_scalar((Kind)k, Span{m.data(),m.size()}, pl); }
  printf("codes: "); for(int i=0;i<8;i++) printf("%d:%d ",i,hist[i]); printf("\n"); return 0; }
