// Host check of lightzero_b200/csrc/lz_exact_math.h against libm expf (the function the reference
// tree calls).  usage: check_expf <stride>   (stride 1 = every float in [-104, +0], ~8 s)
#include <stdio.h>
#include <stdlib.h>
#include "../../lightzero_b200/csrc/lz_exact_math.h"
static inline uint32_t asu32(float f){uint32_t u;memcpy(&u,&f,4);return u;}
static inline float asf(uint32_t u){float f;memcpy(&f,&u,4);return f;}
int main(int argc,char**argv){
  uint32_t stride = argc>1 ? (uint32_t)atoi(argv[1]) : 1;
  unsigned long long bad=0, tot=0;
  // negative floats: 0x80000000 (-0) .. 0xc2d00000 (-104)
  for(uint64_t u=0x80000000ull; u<=0xc2d00000ull; u+=stride){
    float x=asf((uint32_t)u); float ref=expf(x), mine=lz_expf_exact(x);
    if(asu32(ref)!=asu32(mine)){ if(bad<10) printf("mismatch x=%a ref=%a mine=%a\n",x,ref,mine); bad++; }
    tot++;
  }
  float specials[]={-0x1.f8cbb2p+5f,0.0f,-0.0f,-150.0f,-1e30f,-__builtin_huge_valf(),1.0f,0.5f,10.0f,88.0f};
  for(unsigned i=0;i<sizeof(specials)/4;i++){ float x=specials[i]; float ref=expf(x), mine=lz_expf_exact(x);
    if(asu32(ref)!=asu32(mine)){ printf("special mismatch x=%a ref=%a mine=%a\n",x,ref,mine); bad++; } tot++; }
  printf("checked %llu mismatches %llu\n",tot,bad);
  return bad?1:0;
}
