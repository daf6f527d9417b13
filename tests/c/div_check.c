// Checker for the shared-reciprocal division used by the Lennard-Jones kernel (csrc/objectives.cu: div_by):
// with y = RN(1/b), q = RN(a*y), rem = fma(-b, q, a), RN(q + rem*y) must equal the IEEE quotient a/b bit for bit
// (Markstein).  Random operands plus adversarial divisors whose significand is all ones / nearly all ones.
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
static uint64_t s[2] = {0x9E3779B97F4A7C15ull, 0xD1B54A32D192ED03ull};
static inline uint64_t rnd(void){ uint64_t s1=s[0], s0=s[1]; s[0]=s0; s1^=s1<<23; s[1]=s1^s0^(s1>>18)^(s0>>5); return s[1]+s0; }
static inline double mk(uint64_t m, int e){ uint64_t bits=((uint64_t)(1023+e)<<52)|(m&0xFFFFFFFFFFFFFull); double d; memcpy(&d,&bits,8); return d; }
int main(){
  long bad=0, badones=0, n=40000000L;
  for(long i=0;i<n;i++){
    double a=mk(rnd(), (int)(rnd()%40)-20); if(rnd()&1) a=-a;
    double b=mk(rnd(), (int)(rnd()%40)-20);
    if((i&1023)==0){ uint64_t m=0xFFFFFFFFFFFFFull; if(i&1024) m^=(rnd()&7); b=mk(m,(int)(rnd()%10)); }
    double y=1.0/b; double q=a*y; double r=fma(-b,q,a); double q2=fma(r,y,q);
    double t=a/b;
    if(q2!=t){ bad++; uint64_t bb; memcpy(&bb,&b,8); if((bb&0xFFFFFFFFFFFFFull)==0xFFFFFFFFFFFFFull) badones++; if(bad<5) printf("a=%a b=%a got=%a want=%a\n",a,b,q2,t); }
  }
  printf("cases=%ld mismatches=%ld (all-ones b: %ld)\n", n, bad, badones);
  return 0;
}
