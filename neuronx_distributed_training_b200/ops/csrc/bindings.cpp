// Python bindings for the sm_100a extension.
#include <torch/extension.h>
#include <vector>

void gemm(const at::Tensor& a, const at::Tensor& b, at::Tensor& out, bool a_kmajor, bool b_kmajor, bool accumulate);
void gemm_set_2cta(bool on);
void gemm_2cta(const at::Tensor& a, const at::Tensor& b, at::Tensor& out, bool a_kmajor, bool b_kmajor, bool accumulate);
std::vector<at::Tensor> rmsnorm_fwd(const at::Tensor& x, const at::Tensor& w, double eps);
std::vector<at::Tensor> rmsnorm_bwd(const at::Tensor& dy, const at::Tensor& x, const at::Tensor& w, const at::Tensor& rstd);
std::vector<at::Tensor> rope_fwd(const at::Tensor& q, const at::Tensor& k, const at::Tensor& cos, const at::Tensor& sin, bool conj);
at::Tensor swiglu_fwd(const at::Tensor& gu);
at::Tensor swiglu_bwd(const at::Tensor& dout, const at::Tensor& gu);
void adamw_flat(at::Tensor master, at::Tensor param, const at::Tensor& grad, at::Tensor m, at::Tensor v,
                const at::Tensor& decay_mask, double lr, double b1, double b2, double eps, double wd, double bc1,
                double bc2, const at::Tensor& grad_scale, bool sr, int64_t seed);
at::Tensor sq_norm(const at::Tensor& x);
std::vector<at::Tensor> ce_stats(const at::Tensor& logits, const at::Tensor& labels, int64_t vocab_start);
void ce_grad_(at::Tensor logits, const at::Tensor& labels, int64_t vocab_start, const at::Tensor& gmax, const at::Tensor& gsum,
              int64_t ignore_index);
std::vector<at::Tensor> qkv_rope_split_fwd(const at::Tensor& packed, const at::Tensor& cos, const at::Tensor& sin,
                                           int64_t B, int64_t S, int64_t Hq, int64_t Hk);
at::Tensor qkv_rope_split_bwd(const at::Tensor& dq, const at::Tensor& dk, const at::Tensor& dv, const at::Tensor& cos,
                              const at::Tensor& sin);
std::vector<at::Tensor> flash_attn_fwd1(const at::Tensor& q, const at::Tensor& k, const at::Tensor& v, bool causal,
                                        int64_t window, double scale);
std::vector<at::Tensor> flash_attn_fwd(const at::Tensor& q, const at::Tensor& k, const at::Tensor& v, bool causal,
                                       int64_t window, double scale, int64_t q_offset);
void ag_gemm(at::Tensor xg, const at::Tensor& x_local, const at::Tensor& b, at::Tensor out, const at::Tensor& other,
             std::vector<int64_t> xg_peer, std::vector<int64_t> flags_peer, int64_t hs_local_ptr,
             std::vector<int64_t> hs_peer, int64_t epoch, int64_t rank, int64_t tp, int64_t mode, bool b_kmajor,
             bool accumulate, int64_t num_comm_ctas, std::vector<int64_t> x_peer, bool push, int64_t epoch_ptr);
void ag_only(at::Tensor xg, const at::Tensor& x_local, std::vector<int64_t> xg_peer, std::vector<int64_t> flags_peer,
             int64_t hs_local_ptr, std::vector<int64_t> hs_peer, int64_t epoch, int64_t rank, int64_t tp,
             int64_t num_comm_ctas, std::vector<int64_t> x_peer, bool push, int64_t epoch_ptr);
void fused_tp_set_debug(at::Tensor buf);
at::Tensor zero1_reduce(std::vector<int64_t> grads, std::vector<int64_t> params, std::vector<int64_t> hs, int64_t rank,
                        int64_t epoch, at::Tensor gshard, int64_t shard_lo, bool grads_fp32, double scale,
                        const at::Tensor& count_mask);
void zero1_adam_push(std::vector<int64_t> grads, std::vector<int64_t> params, std::vector<int64_t> hs, int64_t rank,
                     int64_t epoch, at::Tensor master, at::Tensor m, at::Tensor v, const at::Tensor& gshard, int64_t shard_lo,
                     std::vector<int64_t> seg_lo, std::vector<int64_t> seg_hi, std::vector<double> seg_lr,
                     std::vector<double> seg_wd, double b1, double b2, double eps, double bc1, double bc2,
                     const at::Tensor& clip, bool sr, int64_t seed, at::Tensor done_ctr);
void gemm_rs(const at::Tensor& a, const at::Tensor& b, at::Tensor out, std::vector<int64_t> stage_peer,
             std::vector<int64_t> counter_peer, int64_t hs_local_ptr, std::vector<int64_t> hs_peer, int64_t epoch,
             int64_t rank, int64_t tp, bool b_kmajor, int64_t epoch_ptr);
void tp_barrier(int64_t hs_local_ptr, std::vector<int64_t> hs_peer, int64_t rank, int64_t tp, int64_t epoch_ptr);
void grouped_gemm(const at::Tensor& a, const at::Tensor& b, at::Tensor& out, const at::Tensor& block_expert, bool b_kmajor);
void grouped_wgrad(const at::Tensor& x, const at::Tensor& dy, at::Tensor& out, const at::Tensor& offsets, bool accumulate);
std::vector<at::Tensor> layernorm_fwd(const at::Tensor& x, const at::Tensor& w, const c10::optional<at::Tensor>& b,
                                      double eps, bool one_plus);
std::vector<at::Tensor> layernorm_bwd(const at::Tensor& dy, const at::Tensor& x, const at::Tensor& w,
                                      const at::Tensor& mean, const at::Tensor& rstd, bool one_plus);
at::Tensor bias_act_fwd(const at::Tensor& x, const c10::optional<at::Tensor>& bias, int64_t mode);
at::Tensor bias_act_bwd(const at::Tensor& dout, const at::Tensor& x, const c10::optional<at::Tensor>& bias, int64_t mode);
at::Tensor scale_mask_softmax_fwd(const at::Tensor& x, double scale, bool causal, int64_t window, int64_t q_offset);
at::Tensor scale_mask_softmax_bwd(const at::Tensor& dy, const at::Tensor& y, double scale);
std::vector<at::Tensor> bias_dropout_add_fwd(const at::Tensor& x, const c10::optional<at::Tensor>& bias,
                                             const at::Tensor& residual, double p);
at::Tensor dropout_bwd(const at::Tensor& dout, const at::Tensor& mask, double p);
at::Tensor embedding_fwd(const at::Tensor& ids, const at::Tensor& w, int64_t vocab_start, bool seq_first);
void embedding_bwd_(at::Tensor dw, const at::Tensor& ids, const at::Tensor& g, int64_t vocab_start, bool seq_first);
std::vector<at::Tensor> flash_attn_bwd(const at::Tensor& dout, const at::Tensor& q, const at::Tensor& k, const at::Tensor& v,
                                       const at::Tensor& out, const at::Tensor& lse, bool causal, int64_t window, double scale,
                                       int64_t q_offset);
void gemm_mxfp8(const at::Tensor& a_q, const at::Tensor& sfa, const at::Tensor& b_q, const at::Tensor& sfb, at::Tensor& out,
                bool accumulate);
std::vector<at::Tensor> mx_quantize_rows(const at::Tensor& x);
std::vector<at::Tensor> mx_quantize_cols_transposed(const at::Tensor& x);
void gemm_set_cta_limit(int64_t n);
int64_t gemm_get_cta_limit();
void flash_fwd2_set_debug(const at::Tensor& t);
void flash_bwd2_set_debug(const at::Tensor& t);
std::vector<at::Tensor> fp8_quantize_rows(const at::Tensor& x);
std::vector<at::Tensor> fp8_quantize_cols_transposed(const at::Tensor& x);
void gemm_fp8(const at::Tensor& a_q, const at::Tensor& a_scale, const at::Tensor& b_q, const at::Tensor& b_scale,
              at::Tensor& out, bool accumulate);

at::Tensor rmsnorm_allgather_mc(const at::Tensor& x, const at::Tensor& w, double eps, int64_t mc_out_ptr, int64_t rank,
                                int64_t world, int64_t mc_flag_ptr, int64_t local_flag_ptr, int64_t counter_before);
int64_t nvls_grid_for_rows(int64_t m);
void nvls_allreduce(const at::Tensor& in, at::Tensor& out, int64_t local_region_ptr, int64_t mc_region_ptr, int64_t world,
                    int64_t mc_flag_ptr, int64_t local_flag_ptr, int64_t counter_before);
int64_t nvls_allreduce_arrivals();

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("rmsnorm_allgather_mc", &rmsnorm_allgather_mc,
        "sequence-parallel RMSNorm whose output is multicast (multimem.st over NVSwitch) into every rank's gathered buffer");
  m.def("nvls_grid_for_rows", &nvls_grid_for_rows);
  m.def("nvls_allreduce", &nvls_allreduce, "one-shot fp32 sum over ranks: staging store, one barrier, multimem.ld_reduce of the whole vector");
  m.def("nvls_allreduce_arrivals", &nvls_allreduce_arrivals);
  m.def("gemm", &gemm, "tcgen05 bf16 GEMM: out (+)= op(a) op(b)");
  m.def("gemm_2cta", &gemm_2cta, "tcgen05 cta_group::2 bf16 GEMM: a cluster of two SMs per 256x256 tile");
  m.def("gemm_set_2cta", &gemm_set_2cta, "route gemm() through the SM-pair kernel when the shape allows (default on)");
  m.def("gemm_set_cta_limit", &gemm_set_cta_limit, "cap the persistent GEMM grid (0 = one CTA per SM): leaves SMs to concurrent collectives");
  m.def("gemm_get_cta_limit", &gemm_get_cta_limit);
  m.def("gemm_mxfp8", &gemm_mxfp8,
        "MX block-scaled e4m3 GEMM: tcgen05.mma.kind::mxf8f6f4.block_scale, 1x32 E8M0 scale factors staged in tensor memory");
  m.def("mx_quantize_rows", &mx_quantize_rows, "bf16 [R,C] -> e4m3 + packed E8M0 1x32 block scales along C");
  m.def("mx_quantize_cols_transposed", &mx_quantize_cols_transposed, "bf16 [R,C] -> e4m3 of x^T [C,R] + packed scales along R");
  m.def("gemm_fp8", &gemm_fp8, "e4m3 tcgen05 GEMM (kind::f8f6f4) with per-row scales of both operands");
  m.def("fp8_quantize_rows", &fp8_quantize_rows, "bf16 [R,C] -> e4m3 bytes + one fp32 scale per row");
  m.def("fp8_quantize_cols_transposed", &fp8_quantize_cols_transposed,
        "bf16 [R,C] -> e4m3 bytes of the transpose [C,R] + one fp32 scale per original column");
  m.def("rmsnorm_fwd", &rmsnorm_fwd);
  m.def("rmsnorm_bwd", &rmsnorm_bwd);
  m.def("rope_fwd", &rope_fwd);
  m.def("swiglu_fwd", &swiglu_fwd);
  m.def("swiglu_bwd", &swiglu_bwd);
  m.def("adamw_flat", &adamw_flat);
  m.def("sq_norm", &sq_norm);
  m.def("ce_stats", &ce_stats);
  m.def("ce_grad_", &ce_grad_);
  m.def("qkv_rope_split_fwd", &qkv_rope_split_fwd);
  m.def("qkv_rope_split_bwd", &qkv_rope_split_bwd);
  namespace py = pybind11;
  m.def("flash_attn_fwd", &flash_attn_fwd, "tcgen05 flash attention forward (head_dim 128); q_offset / longer k, v: context parallelism",
        py::arg("q"), py::arg("k"), py::arg("v"), py::arg("causal"), py::arg("window"), py::arg("scale"), py::arg("q_offset") = 0);
  m.def("flash_fwd2_set_debug", &flash_fwd2_set_debug, "in-kernel clock64 timeline of CTA (0,0) of the forward kernel");
  m.def("flash_bwd2_set_debug", &flash_bwd2_set_debug, "in-kernel clock64 timeline of CTA (0,0) of the backward kernel");
  m.def("flash_attn_fwd1", &flash_attn_fwd1, "single-tile forward kernel (any multiple of 128)");
  m.def("flash_attn_bwd", &flash_attn_bwd,
        "tcgen05 flash attention backward: transposed scores, P^T in TMEM, software-pipelined (training path)",
        py::arg("dout"), py::arg("q"), py::arg("k"), py::arg("v"), py::arg("out"), py::arg("lse"), py::arg("causal"),
        py::arg("window"), py::arg("scale"), py::arg("q_offset") = 0);
  m.def("ag_gemm", &ag_gemm, "fused all-gather -> tcgen05 GEMM over peer memory");
  m.def("ag_only", &ag_only);
  m.def("fused_tp_set_debug", &fused_tp_set_debug);
  m.def("tp_barrier", &tp_barrier, "all-peer barrier over the symmetric handshake slots (device-side epoch)");
  m.def("zero1_reduce", &zero1_reduce, "ZeRO-1 gradient reduce-scatter by peer pulls");
  m.def("zero1_adam_push", &zero1_adam_push, "partitioned AdamW + parameter all-gather by peer pushes");
  m.def("gemm_rs", &gemm_rs, "fused tcgen05 GEMM -> reduce-scatter over peer memory");
  m.def("embedding_fwd", &embedding_fwd, "vocab-parallel embedding lookup (optionally sequence-first output)");
  m.def("embedding_bwd_", &embedding_bwd_, "scatter-add of token gradients into the fp32 main_grad rows");
  m.def("layernorm_fwd", &layernorm_fwd);
  m.def("layernorm_bwd", &layernorm_bwd);
  m.def("bias_act_fwd", &bias_act_fwd, "fused bias + activation (gelu/geglu/reglu/swiglu/relu/squared-relu/silu)");
  m.def("bias_act_bwd", &bias_act_bwd);
  m.def("scale_mask_softmax_fwd", &scale_mask_softmax_fwd, "fused scale + causal/window mask + softmax");
  m.def("scale_mask_softmax_bwd", &scale_mask_softmax_bwd);
  m.def("bias_dropout_add_fwd", &bias_dropout_add_fwd, "residual + dropout(x + bias)");
  m.def("dropout_bwd", &dropout_bwd);
  m.def("grouped_gemm", &grouped_gemm, "per-expert tcgen05 GEMM over 128-row-padded token groups (dropless MoE)");
  m.def("grouped_wgrad", &grouped_wgrad, "per-expert weight-gradient GEMM with device-side row ranges");
}
