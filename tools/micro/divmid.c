// host check of the scaled-quotient division: q = RN(a/b) for tiny a (denormal or zero quotients included)
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <math.h>
#include <string.h>
static inline uint64_t bits(double x){uint64_t u;memcpy(&u,&x,8);return u;}
static inline double frombits(uint64_t u){double x;memcpy(&x,&u,8);return x;}
static uint64_t s=88172645463325252ull;
static inline uint64_t rnd(){s^=s<<13;s^=s>>7;s^=s<<17;return s;}
// reciprocal as the device computes it: some seed + two Newton steps (here: exact 1/b perturbed, then refined)
static double make_r(double b,int perturb){
    double r0=(float)(1.0/b);           // ~24-bit seed
    double t=fma(r0,-b,1.0); t=fma(t,t,t);
    double r1=fma(r0,t,r0);
    double t2=fma(r1,-b,1.0);
    double r=fma(r1,t2,r1);
    if(perturb) r=frombits(bits(r)+perturb);
    return r;
}
static double div_mid(double a,double b,double r,int*slow){
    if(a==0.0) return r*a;
    if(!(fabs(a)<0x1p-800)){*slow=1;return a/b;}
    const double a2=a*0x1p600;
    double q0=r*a2; double e=fma(q0,-b,a2); double q2=fma(r,e,q0);
    double qd=q2*0x1p-600;
    if(fabs(q2)>=0x1p-422) return qd;
    double back=qd*0x1p600, diff=q2-back;
    if(fabs(diff)==0x1p-475){
        double e2=fma(-q2,b,a2);
        if(e2!=0.0){
            int up=((e2>0)==(b>0));           // x2 > q2
            int incmag=(up==(q2>0));
            double q2n=frombits(bits(q2)+(incmag?1:-1));
            qd=q2n*0x1p-600;
        }
    }
    return qd;
}
int main(int argc,char**argv){
    long N=argc>1?atol(argv[1]):20000000; long bad=0,ties=0,slow=0,den=0;
    double bs[]={1.0/(4096.0*4096.0), 1.0/(400.0*400.0),(10.0/400)*(10.0/400),(3.0/400)*(3.0/400), -4.0, -1.3333333e-3, 0.3e-6, 7.77e5, 1.0, 3.0, -0x1.fffffffffffffp-30, 0x1.0000000000001p40};
    int nb=sizeof(bs)/sizeof(bs[0]);
    for(int pb=-1;pb<=1;++pb)
    for(int ib=0;ib<nb;++ib){
        double b=bs[ib]; double r=make_r(b,pb);
        for(long n=0;n<N;++n){
            double a; int mode=n%4;
            if(mode==0){ // random tiny exponent
                int ex=-1074+(int)(rnd()%300); a=ldexp((double)(rnd()>>11)/9007199254740992.0+1.0,ex); }
            else if(mode==1){ // denormal a
                a=frombits(rnd()%(1ull<<52)); }
            else if(mode==2){ // near-tie: a ~ (k+1/2)*2^-1074 * b
                uint64_t k=rnd()%(1ull<<(rnd()%52+1)); double m=ldexp((double)k+0.5,-1074+300); // scaled by 2^300
                double am=m*b; // rounded
                am=frombits(bits(am)+(int)(rnd()%5)-2);
                a=am*0x1p-300; if(a==0) a=ldexp(1.0,-1074); }
            else { int ex=-1000+(int)(rnd()%200); a=ldexp((double)(rnd()>>11)/9007199254740992.0+1.0,ex); }
            if(rnd()&1) a=-a;
            int sl=0; double q=div_mid(a,b,r,&sl); double ref=a/b;
            slow+=sl; if(fabs(ref)<0x1p-1022) den++;
            if(bits(q)!=bits(ref)){ if(bad<10) printf("BAD a=%a b=%a q=%a ref=%a\n",a,b,q,ref); bad++; }
        }
    }
    printf("bad=%ld slow=%ld denormal_results=%ld\n",bad,slow,den);
    return bad!=0;
}
