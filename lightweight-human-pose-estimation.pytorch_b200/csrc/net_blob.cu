// lwp_net: the whole network behind three C calls -- load a self-contained blob, run it, read the outputs -- so a host
// that is not Python can use the forward without re-implementing the module walk, BatchNorm folding and weight packing
// of the Python mirror (lwpose_b200/engine.py).  The blob is written once by Plan.export_blob() from a reference
// state_dict (it is also the persistent pre-folded / pre-packed weight cache: SURVEY.md section 8 row f4) and holds
//   header   "LWPB", version, dtype, n, H, W, #tensors, #ops, #compute ops
//   tensors  per tensor: bytes, kind (0 activation, 1 activation zero-initialised, 2 constant followed by its data)
//   ops      per op: function id (the lwp_plan_add_* entry point), #args, then typed arguments
//            (0 int64 | 1 double | 2 tensor id + byte offset | 3 NULL pointer | 4 double[3])
//   outputs  #stages, #nchw, tensor ids of the float32 head buffers [n*h*w][64] per stage and of the NCHW tensors
// Replaces PoseEstimationWithMobileNet.__init__ + load_state + forward of the reference
// (models/with_mobilenet.py:89-123, modules/load_state.py:4-15) for a fixed (dtype, batch, height, width).
#include "common.cuh"

#include <vector>

struct lwp_net {
  lwp_plan *plan = nullptr;
  int dtype = 0, n = 0, H = 0, W = 0, compute_ops = 0;
  std::vector<void *> tensors;
  std::vector<size_t> bytes;
  std::vector<int> heads, nchw;
};

namespace {

struct Reader {
  const unsigned char *p, *end;
  bool ok = true;
  template <typename T> T get() {
    T v{};
    if (p + sizeof(T) > end) { ok = false; return v; }
    memcpy(&v, p, sizeof(T));
    p += sizeof(T);
    return v;
  }
  const unsigned char *skip(size_t n) {
    if (n > (size_t)(end - p)) { ok = false; return nullptr; }
    const unsigned char *q = p;
    p += n;
    return q;
  }
};

struct Arg {
  int type = 0;
  long long i = 0;
  double d = 0;
  void *ptr = nullptr;
  double d3[3] = {0, 0, 0};
};

}  // namespace

extern "C" void lwp_net_destroy(lwp_net *net) {
  if (net == nullptr) return;
  if (net->plan) lwp_plan_destroy(net->plan);
  for (void *t : net->tensors) if (t) cudaFree(t);
  delete net;
}

extern "C" int lwp_net_load(const void *blob, size_t bytes, lwp_net **out) {
  LWP_REQUIRE(blob != nullptr && out != nullptr, "lwp_net_load: null pointer");
  Reader r{(const unsigned char *)blob, (const unsigned char *)blob + bytes};
  const unsigned char *magic = r.skip(4);
  const uint32_t version = r.get<uint32_t>();
  LWP_REQUIRE(r.ok && memcmp(magic, "LWPB", 4) == 0 && version == 1, "lwp_net_load: not a version-1 LWPB blob");
  lwp_net *net = new lwp_net();
  net->dtype = r.get<int>(); net->n = r.get<int>(); net->H = r.get<int>(); net->W = r.get<int>();
  const int n_tensors = r.get<int>(), n_ops = r.get<int>();
  net->compute_ops = r.get<int>();
  int rc = LWP_OK;
  auto fail = [&](int code) { lwp_net_destroy(net); return code; };
  if (!r.ok || n_tensors < 0 || n_ops < 0) { lwp::set_error("lwp_net_load: truncated header"); return fail(LWP_EINVAL); }
  for (int t = 0; t < n_tensors; ++t) {
    const uint64_t nb = r.get<uint64_t>();
    const int kind = r.get<int>();
    if (!r.ok) { lwp::set_error("lwp_net_load: truncated tensor table"); return fail(LWP_EINVAL); }
    void *d = nullptr;
    if (nb > 0) {
      cudaError_t e = cudaMalloc(&d, nb);
      if (e != cudaSuccess) { lwp::set_error("lwp_net_load: cudaMalloc(%llu): %s", (unsigned long long)nb, cudaGetErrorString(e)); return fail(LWP_ECUDA); }
    }
    net->tensors.push_back(d);
    net->bytes.push_back((size_t)nb);
    if (kind == 2) {
      const unsigned char *src = r.skip((size_t)((nb + 15) / 16 * 16));
      if (!r.ok) { lwp::set_error("lwp_net_load: truncated constant data"); return fail(LWP_EINVAL); }
      if (nb > 0 && cudaMemcpy(d, src, nb, cudaMemcpyHostToDevice) != cudaSuccess) { lwp::set_error("lwp_net_load: cudaMemcpy failed"); return fail(LWP_ECUDA); }
    } else if (nb > 0) {   // activation buffers start from zero (kind 1 relies on it: the pad channels of the concat / head buffers)
      if (cudaMemset(d, 0, nb) != cudaSuccess) { lwp::set_error("lwp_net_load: cudaMemset failed"); return fail(LWP_ECUDA); }
    }
  }
  rc = lwp_plan_create(net->dtype, &net->plan);
  if (rc != LWP_OK) return fail(rc);
  for (int o = 0; o < n_ops; ++o) {
    const int fn = r.get<int>(), na = r.get<int>();
    if (!r.ok || na < 0 || na > 32) { lwp::set_error("lwp_net_load: truncated op list"); return fail(LWP_EINVAL); }
    Arg a[32];
    for (int k = 0; k < na; ++k) {
      a[k].type = r.get<int>();
      switch (a[k].type) {
        case 0: a[k].i = r.get<long long>(); break;
        case 1: a[k].d = r.get<double>(); break;
        case 2: {
          const int id = r.get<int>();
          const uint64_t off = r.get<uint64_t>();
          if (!r.ok || id < 0 || id >= n_tensors || off >= net->bytes[id]) { lwp::set_error("lwp_net_load: bad tensor reference"); return fail(LWP_EINVAL); }
          a[k].ptr = (char *)net->tensors[id] + off;
          break;
        }
        case 3: a[k].ptr = nullptr; break;
        case 4: a[k].d3[0] = r.get<double>(); a[k].d3[1] = r.get<double>(); a[k].d3[2] = r.get<double>(); break;
        default: lwp::set_error("lwp_net_load: unknown argument type %d", a[k].type); return fail(LWP_EINVAL);
      }
    }
    if (!r.ok) { lwp::set_error("lwp_net_load: truncated op arguments"); return fail(LWP_EINVAL); }
#define P(k) a[k].ptr
#define I(k) ((int)a[k].i)
#define FP(k) ((const float *)a[k].ptr)
    lwp_plan *pl = net->plan;
    switch (fn) {
      case 0: if (na != 7) { rc = LWP_EINVAL; break; }
        rc = lwp_plan_add_stem(pl, FP(0), FP(1), FP(2), P(3), I(4), I(5), I(6)); break;
      case 1: if (na != 9) { rc = LWP_EINVAL; break; }
        rc = lwp_plan_add_stem_u8(pl, FP(0), FP(1), FP(2), P(3), I(4), I(5), I(6), a[7].d3, a[8].d); break;
      case 2: if (na != 12) { rc = LWP_EINVAL; break; }
        rc = lwp_plan_add_depthwise(pl, P(0), P(1), FP(2), FP(3), FP(4), I(5), I(6), I(7), I(8), I(9), I(10), I(11)); break;
      case 3: if (na != 19) { rc = LWP_EINVAL; break; }
        rc = lwp_plan_add_conv_gemm(pl, P(0), I(1), P(2), FP(3), FP(4), P(5), I(6), P(7), I(8), (float *)P(9), I(10), I(11), I(12),
                                    I(13), I(14), I(15), I(16), I(17), I(18)); break;
      case 4: if (na != 19) { rc = LWP_EINVAL; break; }
        rc = lwp_plan_add_dwpw(pl, P(0), FP(1), FP(2), FP(3), I(4), I(5), P(6), FP(7), FP(8), I(9), P(10), I(11), P(12), I(13), I(14),
                               I(15), I(16), I(17), I(18)); break;
      case 5: if (na != 19) { rc = LWP_EINVAL; break; }
        rc = lwp_plan_add_sepconv(pl, P(0), FP(1), FP(2), FP(3), I(4), I(5), P(6), FP(7), FP(8), I(9), P(10), I(11), P(12), I(13),
                                  I(14), I(15), I(16), I(17), I(18)); break;
      case 6: if (na != 15) { rc = LWP_EINVAL; break; }
        rc = lwp_plan_add_heads_fused(pl, P(0), I(1), P(2), FP(3), FP(4), I(5), P(6), FP(7), FP(8), P(9), I(10), (float *)P(11), I(12),
                                      I(13), I(14)); break;
      case 7: if (na != 9) { rc = LWP_EINVAL; break; }
        rc = lwp_plan_add_nhwc_to_nchw(pl, P(0), I(1), I(2), I(3), I(4), (float *)P(5), I(6), I(7), I(8)); break;
      case 8: if (na != 19) { rc = LWP_EINVAL; break; }
        rc = lwp_plan_add_frontend(pl, FP(0), FP(1), FP(2), FP(3), FP(4), FP(5), P(6), FP(7), FP(8), FP(9), FP(10), FP(11), P(12),
                                   I(13), I(14), I(15), I(16), a[17].d3, a[18].d); break;
      case 9: if (na != 19) { rc = LWP_EINVAL; break; }
        rc = lwp_plan_add_conv3x3_pw(pl, P(0), I(1), P(2), FP(3), FP(4), P(5), I(6), I(7), P(8), FP(9), FP(10), I(11), P(12), I(13),
                                     I(14), I(15), I(16), I(17), I(18)); break;
      default: rc = LWP_EINVAL;
    }
#undef P
#undef I
#undef FP
    if (rc == LWP_EINVAL && lwp_last_error()[0] == 0) lwp::set_error("lwp_net_load: op %d: bad function id / argument count", o);
    if (rc != LWP_OK) return fail(rc);
  }
  const int n_heads = r.get<int>(), n_nchw = r.get<int>();
  if (!r.ok || n_heads < 1 || n_nchw < 0) { lwp::set_error("lwp_net_load: truncated output table"); return fail(LWP_EINVAL); }
  for (int k = 0; k < n_heads + n_nchw; ++k) {
    const int id = r.get<int>();
    if (!r.ok || id < 0 || id >= n_tensors) { lwp::set_error("lwp_net_load: bad output tensor id"); return fail(LWP_EINVAL); }
    (k < n_heads ? net->heads : net->nchw).push_back(id);
  }
  *out = net;
  return LWP_OK;
}

extern "C" int lwp_net_info(const lwp_net *net, int *dtype, int *n, int *H, int *W, int *n_stages) {
  LWP_REQUIRE(net != nullptr, "lwp_net_info: null net");
  if (dtype) *dtype = net->dtype;
  if (n) *n = net->n;
  if (H) *H = net->H;
  if (W) *W = net->W;
  if (n_stages) *n_stages = (int)net->heads.size();
  return LWP_OK;
}

extern "C" int lwp_net_forward(lwp_net *net, const void *x, int with_nchw, void *stream) {
  LWP_REQUIRE(net != nullptr && x != nullptr, "lwp_net_forward: null pointer");
  return lwp_plan_run_range(net->plan, x, 0, with_nchw ? lwp_plan_num_ops(net->plan) : net->compute_ops, stream);
}

extern "C" int lwp_net_heads(const lwp_net *net, int stage, const float **heads, int *ld) {
  LWP_REQUIRE(net != nullptr && heads != nullptr, "lwp_net_heads: null pointer");
  const int ns = (int)net->heads.size();
  if (stage < 0) stage += ns;
  LWP_REQUIRE(stage >= 0 && stage < ns, "lwp_net_heads: stage out of range");
  *heads = (const float *)net->tensors[net->heads[stage]];
  if (ld) *ld = 64;
  return LWP_OK;
}

extern "C" int lwp_net_output_nchw(const lwp_net *net, int index, const float **out) {
  LWP_REQUIRE(net != nullptr && out != nullptr, "lwp_net_output_nchw: null pointer");
  LWP_REQUIRE(index >= 0 && index < (int)net->nchw.size(), "lwp_net_output_nchw: index out of range");
  *out = (const float *)net->tensors[net->nchw[index]];
  return LWP_OK;
}
