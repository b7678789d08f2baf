// Closed-form window index maps shared by device kernels and host code.
//
// Restates (without materialising anything) the composition the reference performs with tensors:
//   F.pad to (Hp, Wp)  ->  torch.roll(-shift, -shift)  ->  window_partition      (newcrf_layers.py:212-233, :30-42)
// and its inverse window_reverse -> torch.roll(+shift) -> crop (:239-249, :45-59), plus the region ids the
// shifted-window attention mask is built from (:332-350).
#pragma once
#include <stdint.h>

namespace crf {

struct WindowGeom {
  int H, W;      // feature map
  int ws;        // window size (7)
  int shift;     // 0 or ws/2
  int Hp, Wp;    // padded to multiples of ws
  int nWw, nW;   // windows per row / per image
  __host__ __device__ WindowGeom() {}
  __host__ __device__ WindowGeom(int H_, int W_, int ws_, int shift_) : H(H_), W(W_), ws(ws_), shift(shift_) {
    Hp = (H + ws - 1) / ws * ws;
    Wp = (W + ws - 1) / ws * ws;
    nWw = Wp / ws;
    nW = (Hp / ws) * nWw;
  }
  // Token (h*W + w) feeding position p (0..ws*ws-1) of window `win` of an image, or -1 for a zero-pad position.
  __host__ __device__ int source(int win, int p) const {
    const int wh = win / nWw, ww = win - wh * nWw;
    const int i = p / ws, j = p - i * ws;
    int h = wh * ws + i + shift;
    int w = ww * ws + j + shift;
    if (h >= Hp) h -= Hp;
    if (w >= Wp) w -= Wp;
    return (h < H && w < W) ? h * W + w : -1;
  }
  // Region id (0..8) of position p of window `win` on the shifted, padded grid (only meaningful if shift > 0).
  __host__ __device__ int region(int win, int p) const {
    const int wh = win / nWw, ww = win - wh * nWw;
    const int i = p / ws, j = p - i * ws;
    const int h = wh * ws + i, w = ww * ws + j;
    const int rh = h < Hp - ws ? 0 : (h < Hp - shift ? 1 : 2);
    const int rw = w < Wp - ws ? 0 : (w < Wp - shift ? 1 : 2);
    return rh * 3 + rw;
  }
};

}  // namespace crf
