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


SRC3 = r"""
#include "fft2.cuh"
#include <cstdio>
#include <cmath>
#include <vector>
#include <complex>
using namespace fftk;
// host emulation of the row-major c2r x pass (fft_x_c2r3_v4): two real rows per complex column, Hermitian tangle fused into stage A
template<int N> double test() {
  using X = XRow<N>;
  constexpr int R1 = X::R1, RP = X::RP, YP = X::YP, HC = X::HC;
  std::vector<double> real((size_t)2*LX*N);
  for (int row=0; row<2*LX; ++row) for (int x=0;x<N;++x) real[(size_t)row*N+x] = sin(0.37*x*(row+1)) + 0.25*cos(1.1*x - row) + 0.01*row;
  std::vector<float2> Rw((size_t)2*LX*RP, make_float2(7.f,7.f)), Y((size_t)LX*YP, make_float2(9.f,9.f)), twT(X::NTW);
  for (int row=0; row<2*LX; ++row) for (int k=0;k<HC;++k) {
    std::complex<double> s=0; for (int x=0;x<N;++x){ double a=-2*M_PI*(double)x*k/N; s+=real[(size_t)row*N+x]*std::complex<double>(cos(a),sin(a)); }
    // garbage in the imaginary parts of the k = 0 and k = N/2 bins must be ignored (c2r semantics)
    if (k==0 || 2*k==N) s = {s.real(), 123.0};
    Rw[(size_t)row*RP+k]=make_float2((float)s.real(),(float)s.imag());
  }
  for (int r=1;r<R1;++r) for (int j=0;j<16;++j){ double a=-2*M_PI*(double)(r*j)/N; twT[(r-1)*16+j]=make_float2((float)cos(a),(float)sin(a)); }
  std::vector<float2> regs((size_t)X::NA*16);
  for (int tid=0; tid<X::NA; ++tid) { int c=tid/R1, j=tid-c*R1; float2 v[16];
    c2r_stageA_load<N>(Rw.data()+(size_t)(2*c)*RP, Rw.data()+(size_t)(2*c+1)*RP, j, v); PRadix<16,true>::run(v);
    for (int r=0;r<16;++r) regs[(size_t)tid*16+r]=v[r]; }
  for (int tid=0; tid<X::NA; ++tid) { int c=tid/R1, j=tid-c*R1; float2 v[16]; for (int r=0;r<16;++r) v[r]=regs[(size_t)tid*16+r];
    c2r_stageA_store<N>(Y.data()+(size_t)c*YP, j, v); }
  double maxerr=0, maxv=0;
  for (int tid=0; tid<X::NB; ++tid) { int c=tid>>4, j=tid&15; float2 u[R1];
    c2r_stageB_load<N>(Y.data()+(size_t)c*YP, j, u);
    c2r_stageB_finish<N>(u, twT.data(), j, [&](int r, float2 val){
      const int x=j+16*r;
      const double ea=std::abs(val.x/(double)N-real[(size_t)(2*c)*N+x]), eb=std::abs(val.y/(double)N-real[(size_t)(2*c+1)*N+x]);
      if (ea>maxerr) maxerr=ea; if (eb>maxerr) maxerr=eb; maxv=1.3; }); }
  printf("N=%d rowmajor c2r rel_err=%.3e\n", N, maxerr/maxv); return maxerr/maxv;
}
int main(){
  double m=0; double e;
#define T(N) e=test<N>(); if(e>m)m=e;
  T(32) T(48) T(64) T(80) T(112) T(128) T(176) T(256) T(304)
  return m < 2e-6 ? 0 : 1;
}
"""


def test_rowmajor_c2r_core_on_host(tmp_path):
    """fft2.cuh XRow helpers (c2r along x with lanes along the sequence, tangle fused into stage A) emulated thread by thread."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cu = tmp_path / "t3.cu"
    cu.write_text(SRC3)
    exe = tmp_path / "t3"
    subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-w", "--expt-relaxed-constexpr", "--extended-lambda", "-gencode", "arch=compute_100a,code=sm_100a", "-I", os.path.join(root, "cubep3m_b200", "csrc"), "-o", str(exe), str(cu)])
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    assert r.stdout.count("rel_err") == 9


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
