"""CPU emulation of the device FFT butterflies (fft_smem.cuh is __host__ __device__): every supported N, both directions,
against a double-precision DFT. Runs without a GPU (compiled for the host with nvcc)."""
import os
import subprocess

SRC = r'''
#include "fft_smem.cuh"
#include <cstdio>
#include <cmath>
#include <vector>
#include <complex>
using namespace fftk;
template<int N> double test() {
  std::vector<float> re0(N*LXP), im0(N*LXP), re1(N*LXP), im1(N*LXP);
  std::vector<float2> tw(N);
  for (int t=0;t<N;++t){ double a=-2*M_PI*t/N; tw[t]=make_float2((float)cos(a),(float)sin(a)); }
  std::vector<std::complex<double>> x(N*LX);
  for (int e=0;e<N;++e) for(int c=0;c<LX;++c) x[e*LX+c]={sin(e*1.3+c)+0.1*c, cos(e*0.7-c)};
  double maxerr=0, maxv=0;
  for (int inv=0; inv<2; ++inv) {
    for (int e=0;e<N;++e) for(int c=0;c<LX;++c){ re0[e*LXP+c]=(float)x[e*LX+c].real(); im0[e*LXP+c]=(float)x[e*LX+c].imag(); }
    if (inv) fft_columns<N,true>(re0.data(),im0.data(),re1.data(),im1.data(),tw.data());
    else fft_columns<N,false>(re0.data(),im0.data(),re1.data(),im1.data(),tw.data());
    float* rr = result_buffer<N>() ? re1.data() : re0.data();
    float* ri = result_buffer<N>() ? im1.data() : im0.data();
    for (int c=0;c<LX;c+=5) for (int k=0;k<N;++k) {
      std::complex<double> s=0; for (int e=0;e<N;++e){ double a=(inv?2:-2)*M_PI*(double)e*k/N; s+=x[e*LX+c]*std::complex<double>(cos(a),sin(a)); }
      double err=std::abs(s-std::complex<double>(rr[k*LXP+c],ri[k*LXP+c])); if(err>maxerr)maxerr=err; if(std::abs(s)>maxv)maxv=std::abs(s);
    }
  }
  // AoS float2 variant used by the strided passes
  std::vector<float2> a0(N*LX), a1(N*LX);
  for (int inv=0; inv<2; ++inv) {
    for (int e=0;e<N;++e) for(int c=0;c<LX;++c) a0[e*LX+c]=make_float2((float)x[e*LX+c].real(),(float)x[e*LX+c].imag());
    if (inv) fft_columns_aos<N,true>(a0.data(),a1.data(),tw.data());
    else fft_columns_aos<N,false>(a0.data(),a1.data(),tw.data());
    float2* r = result_buffer<N>() ? a1.data() : a0.data();
    for (int c=1;c<LX;c+=5) for (int k=0;k<N;++k) {
      std::complex<double> s=0; for (int e=0;e<N;++e){ double a=(inv?2:-2)*M_PI*(double)e*k/N; s+=x[e*LX+c]*std::complex<double>(cos(a),sin(a)); }
      double err=std::abs(s-std::complex<double>(r[k*LX+c].x,r[k*LX+c].y)); if(err>maxerr)maxerr=err;
    }
  }
  printf("N=%d rel_err=%.3e\n", N, maxerr/maxv); return maxerr/maxv;
}
int main(){
  for (int R=1;R<MAXR;++R) for(int t=0;t<R;++t){ double a=-2*M_PI*t/R; h_w[R][t]=make_float2((float)cos(a),(float)sin(a)); }
  double m=0; double e;
#define T(N) e=test<N>(); if(e>m)m=e;
  T(16) T(32) T(48) T(64) T(80) T(112) T(128) T(176) T(256) T(304) T(512) T(560)
  return m < 5e-7 ? 0 : 1;
}
'''


SRC2 = r'''
#include "fft2.cuh"
#include <cstdio>
#include <cmath>
#include <vector>
#include <complex>
using namespace fftk;
// host emulation of the in-place two-stage transform of fft2.cuh: every phase loops over all threads (a barrier separates phases)
template<int N, int PITCH> double test() {
  using P = Plan2<N>;
  constexpr int R0 = P::R0;
  std::vector<float2> buf(N*PITCH), tw(N);
  for (int t=0;t<N;++t){ double a=-2*M_PI*t/N; tw[t]=make_float2((float)cos(a),(float)sin(a)); }
  std::vector<std::complex<double>> x(N*LX);
  for (int e=0;e<N;++e) for(int c=0;c<LX;++c) x[e*LX+c]={sin(e*1.3+c)+0.1*c, cos(e*0.7-c)};
  double maxerr=0, maxv=0;
  for (int inv=0; inv<2; ++inv) {
    for (int e=0;e<N;++e) for(int c=0;c<LX;++c) buf[e*PITCH+c]=make_float2((float)x[e*LX+c].real(),(float)x[e*LX+c].imag());
    std::vector<float2> regs((size_t)P::NT*R0);
    for (int tid=0; tid<P::NA; ++tid) { float2 v[R0]; int col=tid%LX, j=tid/LX;
      if (inv) { stageA_load<N,true,PITCH>(buf.data(),j,col,v); PRadix<R0,true>::run(v); } else { stageA_load<N,false,PITCH>(buf.data(),j,col,v); PRadix<R0,false>::run(v); }
      for (int r=0;r<R0;++r) regs[(size_t)tid*R0+r]=v[r]; }
    for (int tid=0; tid<P::NA; ++tid) { float2 v[R0]; int col=tid%LX, j=tid/LX; for (int r=0;r<R0;++r) v[r]=regs[(size_t)tid*R0+r]; stageA_store<N,PITCH>(buf.data(),j,col,v); }
    std::vector<float2> res(N*LX);
    for (int tid=0; tid<P::NB; ++tid) { int col=tid%LX, j=tid/LX;
      auto emit=[&](int r, float2 val){ res[(j + r*R0)*LX+col]=val; };
      if (inv) stageB<N,true,PITCH>(buf.data(),tw.data(),j,col,emit); else stageB<N,false,PITCH>(buf.data(),tw.data(),j,col,emit); }
    for (int c=0;c<LX;c+=5) for (int k=0;k<N;++k) {
      std::complex<double> s=0; for (int e=0;e<N;++e){ double a=(inv?2:-2)*M_PI*(double)e*k/N; s+=x[e*LX+c]*std::complex<double>(cos(a),sin(a)); }
      double err=std::abs(s-std::complex<double>(res[k*LX+c].x,res[k*LX+c].y)); if(err>maxerr)maxerr=err; if(std::abs(s)>maxv)maxv=std::abs(s);
    }
  }
  printf("N=%d pitch=%d rel_err=%.3e\n", N, PITCH, maxerr/maxv); return maxerr/maxv;
}
int main(){
  for (int R=1;R<MAXR;++R) for(int t=0;t<R;++t){ double a=-2*M_PI*t/R; h_w[R][t]=make_float2((float)cos(a),(float)sin(a)); }
  double m=0; double e;
#define T(N) e=test<N,16>(); if(e>m)m=e; e=test<N,17>(); if(e>m)m=e;
  T(32) T(48) T(64) T(80) T(112) T(128) T(176) T(256) T(304)
  return m < 5e-7 ? 0 : 1;
}
'''


def test_inplace_two_stage_core_on_host(tmp_path):
    """fft2.cuh (packed butterflies, in-place two-stage Stockham) emulated thread by thread on the host."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cu = tmp_path / "t2.cu"
    cu.write_text(SRC2)
    exe = tmp_path / "t2"
    subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-w", "--expt-relaxed-constexpr", "--extended-lambda", "-gencode", "arch=compute_100a,code=sm_100a", "-I", os.path.join(root, "cubep3m_b200", "csrc"), "-o", str(exe), str(cu)])
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    assert r.stdout.count("rel_err") == 18


def test_butterflies_on_host(tmp_path):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cu = tmp_path / "t.cu"
    cu.write_text(SRC)
    exe = tmp_path / "t"
    subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-w", "-I", os.path.join(root, "cubep3m_b200", "csrc"), "-o", str(exe), str(cu)])
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    assert r.stdout.count("rel_err") == 12
