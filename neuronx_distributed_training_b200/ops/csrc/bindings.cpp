// Python bindings for the sm_100a extension.
#include <torch/extension.h>
#include <vector>

void gemm(const at::Tensor& a, const at::Tensor& b, at::Tensor& out, bool a_kmajor, bool b_kmajor, bool accumulate);
std::vector<at::Tensor> rmsnorm_fwd(const at::Tensor& x, const at::Tensor& w, double eps);
std::vector<at::Tensor> rmsnorm_bwd(const at::Tensor& dy, const at::Tensor& x, const at::Tensor& w, const at::Tensor& rstd);
std::vector<at::Tensor> rope_fwd(const at::Tensor& q, const at::Tensor& k, const at::Tensor& cos, const at::Tensor& sin, bool conj);
at::Tensor swiglu_fwd(const at::Tensor& gu);
at::Tensor swiglu_bwd(const at::Tensor& dout, const at::Tensor& gu);
void adamw_flat(at::Tensor master, at::Tensor param, const at::Tensor& grad, at::Tensor m, at::Tensor v,
                const at::Tensor& decay_mask, double lr, double b1, double b2, double eps, double wd, double bc1,
                double bc2, const at::Tensor& grad_scale, bool sr, int64_t seed);
at::Tensor sq_norm(const at::Tensor& x);

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("gemm", &gemm, "tcgen05 bf16 GEMM: out (+)= op(a) op(b)");
  m.def("rmsnorm_fwd", &rmsnorm_fwd);
  m.def("rmsnorm_bwd", &rmsnorm_bwd);
  m.def("rope_fwd", &rope_fwd);
  m.def("swiglu_fwd", &swiglu_fwd);
  m.def("swiglu_bwd", &swiglu_bwd);
  m.def("adamw_flat", &adamw_flat);
  m.def("sq_norm", &sq_norm);
}
