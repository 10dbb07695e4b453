// Column-major owning array of up to four dimensions, first index fastest.
// Same indexing contract as the reference container (reference include/hf_array.h:303-325): the layouts of
// every array on the hot path (disu_upts(upt,ele,field), JGinv(l,m,pt,ele), ...) follow from it, and the device
// mirror keeps exactly this layout, so a download is a flat copy.
#pragma once
#include <vector>
#include <cstddef>
#include <cstring>
#include <algorithm>

template <typename T>
class hf_array
{
public:
  hf_array() { d_[0] = d_[1] = d_[2] = d_[3] = 0; }
  explicit hf_array(int n0, int n1 = 1, int n2 = 1, int n3 = 1) { setup(n0, n1, n2, n3); }

  void setup(int n0, int n1 = 1, int n2 = 1, int n3 = 1)
  {
    d_[0] = n0; d_[1] = n1; d_[2] = n2; d_[3] = n3;
    v_.assign((size_t)n0 * n1 * n2 * n3, T());
  }

  T &operator()(int i0) { return v_[i0]; }
  T &operator()(int i0, int i1) { return v_[i0 + (size_t)d_[0] * i1]; }
  T &operator()(int i0, int i1, int i2) { return v_[i0 + (size_t)d_[0] * (i1 + (size_t)d_[1] * i2)]; }
  T &operator()(int i0, int i1, int i2, int i3) { return v_[i0 + (size_t)d_[0] * (i1 + (size_t)d_[1] * (i2 + (size_t)d_[2] * i3))]; }
  const T &operator()(int i0) const { return v_[i0]; }
  const T &operator()(int i0, int i1) const { return v_[i0 + (size_t)d_[0] * i1]; }
  const T &operator()(int i0, int i1, int i2) const { return v_[i0 + (size_t)d_[0] * (i1 + (size_t)d_[1] * i2)]; }
  const T &operator()(int i0, int i1, int i2, int i3) const { return v_[i0 + (size_t)d_[0] * (i1 + (size_t)d_[1] * (i2 + (size_t)d_[2] * i3))]; }
  T &operator[](size_t i) { return v_[i]; }

  T *get_ptr_cpu() { return v_.empty() ? nullptr : v_.data(); }
  const T *get_ptr_cpu() const { return v_.empty() ? nullptr : v_.data(); }
  T *get_ptr_cpu(int i0, int i1 = 0, int i2 = 0, int i3 = 0) { return &(*this)(i0, i1, i2, i3); }

  int get_dim(int i) const { return d_[i]; }
  size_t size() const { return v_.size(); }
  void initialize_to_zero() { std::fill(v_.begin(), v_.end(), T()); }
  void initialize_to_value(const T val) { std::fill(v_.begin(), v_.end(), val); }

private:
  int d_[4];
  std::vector<T> v_;
};
