/* Test infrastructure: checks the shared-reciprocal division used by the MODE_NORMAL epilogue of the CUDA stencil kernel
 * (peleanalysis_b200/csrc/stencil_tma.cu: div_by) against the plain IEEE division the reference performs
 * (curvature.cpp:498-502, MultiFab::Divide).  y = RN(1/b); q0 = RN(a*y); r = fma(-b, q0, a); q = fma(r, y, q0) must
 * equal a/b bit for bit.  Operands: random and adversarial mantissas (all ones, near 1.0, equal / adjacent mantissas),
 * |a| <= ~|b| as for a vector component over its norm.  usage: markstein_div_check <seed> <count>   (needs -mfma)
 * Round-1 record: 6 seeds x 8e9 pairs + 1.5e9 pairs of a wider-exponent variant, 0 mismatches. */
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
static uint64_t s[2];
static inline uint64_t rnd(){uint64_t a=s[0],b=s[1];s[0]=b;a^=a<<23;s[1]=a^b^(a>>17)^(b>>26);return s[1]+b;}
static inline double mk(uint64_t m,int e){uint64_t u=((uint64_t)(1023+e)<<52)|(m&0xFFFFFFFFFFFFFULL);double d;memcpy(&d,&u,8);return d;}
int main(int argc,char**argv){
  s[0]=0x9E3779B97F4A7C15ULL*(uint64_t)(atoi(argv[1])+1); s[1]=0xD1B54A32D192ED03ULL^(uint64_t)atoi(argv[1]);
  long N=atol(argv[2]);
  long bad1=0,n=0;
  for(long it=0;it<N;++it){
    uint64_t ma=rnd(),mb=rnd();
    uint64_t re=rnd();
    int eb=(int)(re%340)-46;           /* |b| in [1e-14, 1e88] */
    int ea=eb-(int)((re>>20)%60);      /* |a| <= ~|b| (a is a component of the vector whose norm is b), down to 2^-60 relative */
    if(ea<-800) ea=-800;
    switch(it&7){case 0: mb|=0xFFFFFFFFF0000ULL;break; case 1: mb&=0xFFFFULL;break; case 2: ma|=0xFFFFFFFFFF000ULL;break;
      case 3: ma&=0xFFFULL; mb|=0xFFFFFFFFFFF00ULL;break; case 4: ma=mb; break; case 5: ma=mb+1; break; case 6: ma=mb-1; break; default:break;}
    double a=mk(ma,ea), b=-mk(mb,eb);
    if(it&8) a=-a;
    double y=1.0/b;
    double q0=a*y; double r=fma(-b,q0,a); double q1=fma(r,y,q0);
    double t=a/b;
    if(q1!=t){bad1++; if(bad1<5) printf("mismatch a=%a b=%a q1=%a t=%a\n",a,b,q1,t);}
    n++;
  }
  printf("seed %s n=%ld bad=%ld\n",argv[1],n,bad1);
  return 0;
}
