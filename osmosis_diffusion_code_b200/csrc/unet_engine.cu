// Native UNet engine: topology of UNetModel (unet.py:503-695) for the shipped configuration family
// (use_scale_shift_norm, resblock_updown, legacy attention order, fp32, no class conditioning), state_dict
// ingest + weight repack, activation arena planning, and the forward / input-VJP launch programs.
//
// Data layout in HBM (all fp32):
//   weights   : per conv a forward pack Wf[tap][Cin_p/32][Cout_p][32] and a dgrad pack Wd[tap'][Cout_p/32][Cin_p][32]
//               (flipped taps, transposed) - K-block-major, so every TMA B-operand tile is one contiguous run.
//   activations: NHWC.  Every tensor the input-VJP needs (the 101 GroupNorm inputs, qkv of each attention
//               block, per-group mean/rstd) is a persistent arena slot; everything else is scratch.
//   skip concat: `th.cat([h, hs.pop()], dim=1)` (unet.py:739) is never materialised by a copy - the
//               producer of hs[i] writes straight into the channel slice of the concat buffer its output
//               block will read, and the producer of h writes the other slice (views with ld = Ch+Cs).
//               The same aliasing is used for the gradients, so the skip-gradient fan-in is an `+=` in
//               the consumer's epilogue.
#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"

namespace osm {

namespace {

struct ParamInfo {
  std::string name;
  std::vector<int64_t> shape;
  int kind;   // 0 raw copy, 1 conv weight (repack), 2 emb linear weight (row offset), 3 emb linear bias
  int index;  // conv index / raw slot index
  int64_t aux;
};

struct ConvLayer {
  int Cin, Cout, Cin_p, Cout_p, taps;
  float *wf = nullptr, *wd = nullptr, *bias = nullptr;
  void *wf16 = nullptr, *wd16 = nullptr;   // fp16 packs for the fp16-operand kernels: forward needs Cin_p % 64 == 0, dgrad Cout_p % 64 == 0
  bool f16_fwd_ok() const { return Cin_p % 64 == 0; }
  bool f16_dgrad_ok() const { return Cout_p % 64 == 0; }
};

struct RawSlot {
  int64_t numel;
  float* dev = nullptr;
};

enum LayerKind { L_CONV_IN, L_RES, L_ATTN };
struct Layer {
  LayerKind kind;
  int cin, cout, updown;  // res
  int conv1 = -1, conv2 = -1, skip = -1;
  int g1 = -1, b1 = -1, g2 = -1, b2 = -1;  // raw slots
  int emb_off = 0;
  int heads = 0, qkv = -1, proj = -1;  // attn (g1/b1 = norm)
};

enum OpKind { OP_CONV, OP_GN_STATS, OP_GN_APPLY, OP_GN_BWD, OP_ATTN_FWD, OP_ATTN_BWD, OP_LINEAR, OP_GN_FINALIZE, OP_GN_COEF, OP_GN_COEF_FWD,
              OP_GN_COEF_BATCH, OP_GN_COMBINE };
struct Op {
  OpKind kind;
  ConvTcPlan tc;  // holds ConvArgs too
  GnArgs gn;
  float* gn_y = nullptr;
  GnBwdArgs gnb;
  int gnb_apply_only = 0;    // the two means were reduced in the producing conv's epilogue
  int gn_small = 0;          // small tensor: statistics + apply (forward) / both backward passes in ONE launch (norm.cu)
  // fused-statistics helpers
  const float* fin_partial = nullptr; const float* fin_in = nullptr; float* fin_out = nullptr;
  int fin_slots = 0, fin_HW = 0, fin_C = 0, fin_mode = 0;
  float* gn_coef = nullptr;
  bool fin_has_coef = false; GnArgs fin_gn{}; float* fin_coef = nullptr;   // finalize also writes the consumer's forward coefficients
  const GnCoefDesc* cb_table = nullptr; int cb_n = 0;                       // OP_GN_COEF_BATCH
  const float* cmb_a = nullptr; const float* cmb_b = nullptr; float* cmb_out = nullptr;   // OP_GN_COMBINE
  // attention
  const float* at_qkv = nullptr; const float* at_g = nullptr; float* at_out = nullptr;
  int at_L = 0, at_C = 0, at_heads = 0;
  int at_flash = 0;          // 1: fused tcgen05 kernels (attention_flash.cu) through `fa`
  AttnFlashPlan fa;
  // linear
  const float* li_in = nullptr; const float* li_w = nullptr; const float* li_b = nullptr; float* li_out = nullptr;
  int li_ldin = 0, li_ldout = 0, li_K = 0, li_N = 0, li_silu = 0;
  // profiling metadata: algorithmic FLOPs (2*MAC) and minimum HBM bytes of this op, and a shape label
  double flops = 0, bytes = 0;
  int dims[6] = {0, 0, 0, 0, 0, 0};  // conv: H, W, Cin_p, Cout_p, taps, dgrad ; gn: H, W, C ; attn: L, C, heads
};

inline int pad32(int c) { return (c + 31) / 32 * 32; }

}  // namespace

struct Engine {
  osm_unet_config cfg;
  int conv_mode;
  bool use_flash_attention = [] { const char* e = getenv("OSM_ATTN_FLASH"); return e ? atoi(e) != 0 : true; }();
  std::vector<ParamInfo> params;
  std::map<std::string, int> pindex;
  std::vector<ConvLayer> convs;
  std::vector<RawSlot> raws;
  std::vector<Layer> layers;
  std::vector<std::vector<int>> in_blocks, out_blocks;
  std::vector<int> mid_block;
  int conv_in = -1, conv_out = -1, out_g = -1, out_b = -1;
  int te0_w = -1, te0_b = -1, te2_w = -1, te2_b = -1;
  int emb_total = 0;
  int emb_w_slot = -1, emb_b_slot = -1;

  // device storage
  float* wblock = nullptr;
  float* stage = nullptr;  // staging buffer for repack
  int64_t stage_cap = 0;

  // bound plan
  int B = 0, H = 0, W = 0;
  std::vector<Op> fwd, bwd;
  float *xin = nullptr, *yout = nullptr, *e0 = nullptr, *gy = nullptr, *gxin = nullptr;
  int fwd_launches = 0, bwd_launches = 0;
  double flops = 0;
  bool bound = false;

  int add_raw(const std::string& name, std::vector<int64_t> shape, int64_t padded = 0) {
    int64_t n = 1;
    for (auto s : shape) n *= s;
    raws.push_back(RawSlot{padded > n ? padded : n, nullptr});
    params.push_back(ParamInfo{name, shape, 0, (int)raws.size() - 1, n});
    return (int)raws.size() - 1;
  }
  int add_conv(const std::string& prefix, int cin, int cout, int taps, bool conv1d) {
    ConvLayer c;
    // the 4-channel input / 8-channel output are padded to io_pad channels (64 with fp16 operands: one K block / one 64-wide tile)
    c.Cin = cin; c.Cout = cout; c.Cin_p = cin < 32 ? io_pad() : pad32(cin); c.Cout_p = cout < 32 ? io_pad() : pad32(cout); c.taps = taps;
    convs.push_back(c);
    const int idx = (int)convs.size() - 1;
    std::vector<int64_t> shp = conv1d ? std::vector<int64_t>{cout, cin, 1}
                                      : (taps == 9 ? std::vector<int64_t>{cout, cin, 3, 3} : std::vector<int64_t>{cout, cin, 1, 1});
    params.push_back(ParamInfo{prefix + ".weight", shp, 1, idx, 0});
    params.push_back(ParamInfo{prefix + ".bias", {cout}, 4, idx, 0});
    return idx;
  }
  int io_pad() const { return conv_mode == 0 && use_f16 && use_io64 ? 64 : 32; }
  int use_io64 = [] { const char* e = getenv("OSM_IO_PAD64"); return e ? atoi(e) : 1; }();
  int heads_for(int ch) const { return cfg.num_head_channels == -1 ? cfg.num_heads : ch / cfg.num_head_channels; }
  bool attn_at(int ds) const {
    for (int i = 0; i < cfg.num_attention_ds; ++i)
      if (cfg.attention_ds[i] == ds) return true;
    return false;
  }
  int add_res(const std::string& p, int cin, int cout, int updown) {
    Layer l{};
    l.kind = L_RES; l.cin = cin; l.cout = cout; l.updown = updown;
    l.g1 = add_raw(p + ".in_layers.0.weight", {cin});
    l.b1 = add_raw(p + ".in_layers.0.bias", {cin});
    l.conv1 = add_conv(p + ".in_layers.2", cin, cout, 9, false);
    l.emb_off = emb_total;
    params.push_back(ParamInfo{p + ".emb_layers.1.weight", {2 * cout, 4 * cfg.model_channels}, 2, 0, emb_total});
    params.push_back(ParamInfo{p + ".emb_layers.1.bias", {2 * cout}, 3, 0, emb_total});
    emb_total += 2 * cout;
    l.g2 = add_raw(p + ".out_layers.0.weight", {cout});
    l.b2 = add_raw(p + ".out_layers.0.bias", {cout});
    l.conv2 = add_conv(p + ".out_layers.3", cout, cout, 9, false);
    if (cin != cout) l.skip = add_conv(p + ".skip_connection", cin, cout, 1, false);
    layers.push_back(l);
    return (int)layers.size() - 1;
  }
  int add_attn(const std::string& p, int ch) {
    Layer l{};
    l.kind = L_ATTN; l.cin = l.cout = ch; l.heads = heads_for(ch);
    l.g1 = add_raw(p + ".norm.weight", {ch});
    l.b1 = add_raw(p + ".norm.bias", {ch});
    l.qkv = add_conv(p + ".qkv", ch, 3 * ch, 1, true);
    l.proj = add_conv(p + ".proj_out", ch, ch, 1, true);
    layers.push_back(l);
    return (int)layers.size() - 1;
  }

  int build() {
    const int mc = cfg.model_channels, ted = 4 * mc;
    if (mc % 128) return fail(OSM_ERR_INVALID, "model_channels must be a multiple of 128 (GroupNorm32 groups of >= 4 channels)");
    te0_w = add_raw("time_embed.0.weight", {ted, mc}); te0_b = add_raw("time_embed.0.bias", {ted});
    te2_w = add_raw("time_embed.2.weight", {ted, ted}); te2_b = add_raw("time_embed.2.bias", {ted});
    int ch = cfg.channel_mult[0] * mc;
    {
      Layer l{};
      l.kind = L_CONV_IN; l.cin = cfg.in_channels; l.cout = ch;
      l.conv1 = conv_in = add_conv("input_blocks.0.0", cfg.in_channels, ch, 9, false);
      layers.push_back(l);
      in_blocks.push_back({(int)layers.size() - 1});
    }
    std::vector<int> chans{ch};
    int ds = 1;
    for (int level = 0; level < cfg.num_levels; ++level) {
      const int mult = cfg.channel_mult[level];
      for (int r = 0; r < cfg.num_res_blocks; ++r) {
        const std::string p = "input_blocks." + std::to_string(in_blocks.size());
        std::vector<int> blk{add_res(p + ".0", ch, mult * mc, RS_NONE)};
        ch = mult * mc;
        if (attn_at(ds)) blk.push_back(add_attn(p + ".1", ch));
        in_blocks.push_back(blk);
        chans.push_back(ch);
      }
      if (level != cfg.num_levels - 1) {
        const std::string p = "input_blocks." + std::to_string(in_blocks.size());
        in_blocks.push_back({add_res(p + ".0", ch, ch, RS_DOWN)});
        chans.push_back(ch);
        ds *= 2;
      }
    }
    mid_block.push_back(add_res("middle_block.0", ch, ch, RS_NONE));
    mid_block.push_back(add_attn("middle_block.1", ch));
    mid_block.push_back(add_res("middle_block.2", ch, ch, RS_NONE));
    for (int level = cfg.num_levels - 1; level >= 0; --level) {
      const int mult = cfg.channel_mult[level];
      for (int i = 0; i <= cfg.num_res_blocks; ++i) {
        const int ich = chans.back();
        chans.pop_back();
        const std::string p = "output_blocks." + std::to_string(out_blocks.size());
        int j = 0;
        std::vector<int> blk{add_res(p + "." + std::to_string(j++), ch + ich, mc * mult, RS_NONE)};
        ch = mc * mult;
        if (attn_at(ds)) blk.push_back(add_attn(p + "." + std::to_string(j++), ch));
        if (level && i == cfg.num_res_blocks) {
          blk.push_back(add_res(p + "." + std::to_string(j++), ch, ch, RS_UP));
          ds /= 2;
        }
        out_blocks.push_back(blk);
      }
    }
    out_g = add_raw("out.0.weight", {ch});
    out_b = add_raw("out.0.bias", {ch});
    conv_out = add_conv("out.2", ch, cfg.out_channels, 9, false);
    // packed emb-linear storage
    raws.push_back(RawSlot{(int64_t)emb_total * ted, nullptr}); emb_w_slot = (int)raws.size() - 1;
    raws.push_back(RawSlot{(int64_t)emb_total, nullptr});       emb_b_slot = (int)raws.size() - 1;
    for (size_t i = 0; i < params.size(); ++i) pindex[params[i].name] = (int)i;
    return OSM_OK;
  }

  int ensure_storage() {
    if (wblock) return OSM_OK;
    int64_t total = 0;
    auto rnd = [](int64_t n) { return (n + 63) / 64 * 64; };
    for (auto& r : raws) total += rnd(r.numel);
    const bool want16 = conv_mode == 0 && use_f16;
    for (auto& c : convs) {
      total += 2 * rnd((int64_t)c.taps * c.Cout_p * c.Cin_p) + rnd(c.Cout_p);
      if (want16) total += rnd((int64_t)c.taps * c.Cout_p * c.Cin_p);   // two fp16 packs = one fp32 pack's bytes
    }
    OSM_CUDA_CHECK(cudaMalloc(&wblock, total * sizeof(float)));
    OSM_CUDA_CHECK(cudaMemset(wblock, 0, total * sizeof(float)));
    float* p = wblock;
    for (auto& r : raws) { r.dev = p; p += rnd(r.numel); }
    for (auto& c : convs) {
      const int64_t n = rnd((int64_t)c.taps * c.Cout_p * c.Cin_p);
      c.wf = p; p += n; c.wd = p; p += n; c.bias = p; p += rnd(c.Cout_p);
      if (want16) {
        if (c.f16_fwd_ok()) c.wf16 = p;
        if (c.f16_dgrad_ok()) c.wd16 = p + n / 2;
        p += n;
      }
    }
    return OSM_OK;
  }

  int load_param(const char* name, const float* host, int64_t numel, cudaStream_t s) {
    auto it = pindex.find(name);
    if (it == pindex.end()) return fail(OSM_ERR_INVALID, std::string("unknown parameter: ") + name);
    const ParamInfo& pi = params[it->second];
    int64_t n = 1;
    for (auto d : pi.shape) n *= d;
    if (n != numel) return fail(OSM_ERR_INVALID, std::string("size mismatch for ") + name);
    if (int e = ensure_storage()) return e;
    const int ted = 4 * cfg.model_channels;
    switch (pi.kind) {
      case 0:
        OSM_CUDA_CHECK(cudaMemcpyAsync(raws[pi.index].dev, host, n * 4, cudaMemcpyHostToDevice, s));
        break;
      case 2:
        OSM_CUDA_CHECK(cudaMemcpyAsync(raws[emb_w_slot].dev + pi.aux * ted, host, n * 4, cudaMemcpyHostToDevice, s));
        break;
      case 3:
        OSM_CUDA_CHECK(cudaMemcpyAsync(raws[emb_b_slot].dev + pi.aux, host, n * 4, cudaMemcpyHostToDevice, s));
        break;
      case 4:
        OSM_CUDA_CHECK(cudaMemcpyAsync(convs[pi.index].bias, host, n * 4, cudaMemcpyHostToDevice, s));
        break;
      case 1: {
        ConvLayer& c = convs[pi.index];
        if (n > stage_cap) {
          OSM_CUDA_CHECK(cudaStreamSynchronize(s));
          if (stage) cudaFree(stage);
          OSM_CUDA_CHECK(cudaMalloc(&stage, n * 4));
          stage_cap = n;
        }
        OSM_CUDA_CHECK(cudaMemcpyAsync(stage, host, n * 4, cudaMemcpyHostToDevice, s));
        if (int e = pack_conv_weight_launch(stage, c.wf, c.wd, c.Cout, c.Cin, c.Cout_p, c.Cin_p, c.taps, conv_mode == 0, s)) return e;
        if (c.wf16 || c.wd16)
          if (int e = pack_conv_weight_f16_launch(stage, c.wf16, c.wd16, c.Cout, c.Cin, c.Cout_p, c.Cin_p, c.taps, s)) return e;
        // the host buffer may be freed by the caller right after we return
        OSM_CUDA_CHECK(cudaStreamSynchronize(s));
        break;
      }
    }
    if (pi.kind != 1) OSM_CUDA_CHECK(cudaStreamSynchronize(s));
    return OSM_OK;
  }

  // ------------------------------------------------------------------ planning -----------------------
  struct Arena {
    char* base;
    size_t off = 0;
    float* alloc(size_t nfloats) {
      off = (off + 255) / 256 * 256;
      float* p = base ? (float*)(base + off) : nullptr;
      off += nfloats * sizeof(float);
      return p;
    }
  };
  struct Scratch {
    size_t sa = 0, sb = 0, pd = 0, st = 0, sp = 0, sc = 0;
  };

  struct PlanCtx {
    Engine* e;
    Arena ar;
    bool dry;
    Scratch need;
    float *SA = nullptr, *SB = nullptr, *P = nullptr, *D = nullptr, *ST = nullptr, *bstats = nullptr;
    float *SP = nullptr, *SC = nullptr;                        // fused-statistics partials / per-channel coefficients
    float* SK = nullptr;                                       // cluster split-K partial tiles (conv_tc_kernel), this engine's own
    size_t need_sp_alloc = 0;                                  // floats available at SP
    std::map<std::pair<const float*, int>, float*> fused_stats;  // (view pointer, channels) -> [B][32][2] already reduced
    double* partial = nullptr;
    unsigned int* counter = nullptr;
    float* embout = nullptr;
    std::map<const float*, bool> gwritten;
    std::map<const float*, const float*> cat_partner;  // gradient of a whole concat buffer -> its hs-slice key
    struct Rec {  // what the backward of one layer needs; emitted in reverse layer order
      const Layer* l; View x, gx, y, gy; GnArgs gn1, gn2; View h1, qkv; int flash = 0; AttnFlashPlan fa;
    };
    std::vector<Rec> recs;
    std::vector<GnCoefDesc> coef_descs;
    int err = OSM_OK;
  };

  // Request to reduce GroupNorm statistics of the conv's output in its epilogue (conv_epilogue.cuh).
  //   mode 1: forward stats of `out` -> stats_out (mean, rstd)      mode 2: backward means of GroupNorm `gn` -> stats_out
  struct FuseReq {
    int mode = 0;
    float* stats_out = nullptr;
    GnArgs gn{};
  };
  // 0: off.  1: forward statistics only (+2 % on the conv; the stand-alone statistics pass disappears).  2 (default): also
  // the two backward means, reduced in the dgrad conv's epilogue by EIGHT epilogue warps (with four the epilogue became the
  // critical path: +46 % on the conv; with eight +19 %, less than the reduction pass it replaces - measured per step:
  // B=1 17.55 -> 17.23 ms, B=8 89.4 -> 87.6 ms).
  int use_gn_fusion = [] { const char* e = getenv("OSM_GN_FUSE"); return e ? atoi(e) : 2; }();

  // returns true when the statistics request was honoured (only decided in the non-dry pass)
  // xf_coef != null: the conv reads the RAW GroupNorm input x and applies tf32(SiLU?(x a + b)) to its operand in shared
  // memory (halo kernel, conv_tc.cu) - the stand-alone GroupNorm apply pass does not exist for this conv.
  bool emit_conv(PlanCtx& c, std::vector<Op>& ops, int conv_idx, bool dgrad, View x, View out, const float* bias, View res,
                 int res_mode, int accumulate, const FuseReq* fr = nullptr, const float* xf_coef = nullptr, int xf_silu = 0,
                 bool halo = false, bool x16 = false) {
    const ConvLayer& cl = convs[conv_idx];
    ConvArgs a{};
    a.x = x.p; a.ldx = x.ld;
    a.w = dgrad ? cl.wd : cl.wf;
    a.bias = bias;
    a.res = res.p; a.ldr = res.ld; a.res_mode = res_mode;
    a.out = out.p; a.ldo = out.ld; a.accumulate = accumulate;
    a.B = B; a.H = out.H; a.W = out.W;
    a.Cin_p = dgrad ? cl.Cout_p : cl.Cin_p;
    a.Cout_p = dgrad ? cl.Cin_p : cl.Cout_p;
    a.taps = cl.taps;
    a.splitk_ws = c.SK; a.splitk_ws_bytes = c.SK ? SPLITK_WS_FLOATS * sizeof(float) : 0;
    const void* w16 = dgrad ? cl.wd16 : cl.wf16;
    if (x16) {   // x is an fp16 tensor (its producer wrote it that way, see nh16): the tile / persistent / CTA-pair kernels in kind::f16
      a.f16 = 1; a.w = (const float*)w16;
    } else if (w16 && a.taps == 9 && f16_wanted(a.H, a.W, a.Cin_p, a.Cout_p, a.taps)) {
      a.halo = 1; a.f16 = 1; a.w = (const float*)w16; a.xf_coef = xf_coef; a.xf_silu = xf_silu;
    } else if (halo || xf_coef || halo_wanted(a.H, a.W, a.Cin_p, a.Cout_p, a.taps)) { a.halo = 1; a.xf_coef = xf_coef; a.xf_silu = xf_silu; }
    flops_acc += 2.0 * B * out.H * out.W * (double)a.Cin_p * a.Cout_p * a.taps * (dgrad ? 0 : 1);
    if (fr && fr->mode) {  // scratch sizes do not depend on whether the request ends up honoured
      need(c.need.sp, (size_t)B * ((size_t)(out.H + 7) / 8 + 1) * ((size_t)(out.W + 7) / 8 + 1) * 4 * 64);
      need(c.need.sc, (size_t)B * a.Cout_p * 4);
    }
    // backward statistics: the conv's own (a, b, e) table (filled by ONE batched launch at the start of the program)
    float* cbuf = (fr && fr->mode == 2 && use_coef_batch) ? c.ar.alloc((size_t)B * a.Cout_p * 4) : nullptr;
    if (c.dry) { ops.emplace_back(); return false; }
    Op op{};
    op.kind = OP_CONV;
    bool fused = false;
    if (conv_mode == 0) {
      if (int e = conv_tc_plan(a, &op.tc)) c.err = e;
      const int cpg = a.Cout_p / 32;
      if (fr && fr->mode && use_gn_fusion >= fr->mode && c.SP && conv_tc_stats_capable(op.tc) && out.C == a.Cout_p && a.Cout_p % 32 == 0 &&
          (cpg == 4 || cpg == 8 || cpg == 16 || cpg == 32) &&
          (size_t)B * conv_tc_stat_slots(op.tc) * 64 <= c.need_sp_alloc &&
          (fr->mode == 1 || (fr->gn.resample == RS_NONE && fr->gn.C == a.Cout_p && fr->gn.H == out.H && fr->gn.W == out.W))) {
        fused = true;
        ConvArgs& ca = op.tc.a;
        ca.stat_mode = fr->mode; ca.stat_cpg = cpg; ca.stat_partial = c.SP;
        if (fr->mode == 2) {
          ca.stat_x = fr->gn.x; ca.stat_ldx = fr->gn.ldx; ca.stat_silu = fr->gn.silu;
          if (cbuf && (int)c.coef_descs.size() < COEF_TABLE_CAP) {
            ca.stat_coef = cbuf;
            GnCoefDesc d{};
            d.stats = fr->gn.stats; d.gamma = fr->gn.gamma; d.beta = fr->gn.beta; d.ss = fr->gn.scale_shift; d.coef = (float4*)cbuf;
            d.ld_ss = fr->gn.ld_ss; d.C = a.Cout_p;
            c.coef_descs.push_back(d);
          } else {
            ca.stat_coef = c.SC;
            Op k{}; k.kind = OP_GN_COEF; k.gn = fr->gn; k.gn_coef = c.SC; k.bytes = 16.0 * B * a.Cout_p;
            k.dims[0] = out.H; k.dims[1] = out.W; k.dims[2] = a.Cout_p;
            ops.push_back(k);
          }
        }
      }
    } else {
      op.tc.a = a;
    }
    const double px = (double)B * out.H * out.W;
    op.flops = 2.0 * px * a.Cin_p * a.Cout_p * a.taps;
    op.bytes = 4.0 * (px * a.Cin_p + px * a.Cout_p * (1 + (res_mode != RES_NONE) + (accumulate != 0)) +
                      (double)a.taps * a.Cin_p * a.Cout_p);
    // dims[5]: bit 0 = input-gradient conv; bits 1-2 = operand type: 0 TF32, 1 fp16 halo kernel (fp32 input converted in shared memory),
    // 2 fp16 operands read from an fp16 tensor
    op.dims[0] = out.H; op.dims[1] = out.W; op.dims[2] = a.Cin_p; op.dims[3] = a.Cout_p; op.dims[4] = a.taps;
    op.dims[5] = (dgrad ? 1 : 0) + 2 * (a.f16 ? (a.halo ? 1 : 2) : 0);
    ops.push_back(op);
    if (fused) {
      Op f{}; f.kind = OP_GN_FINALIZE;
      f.fin_partial = c.SP; f.fin_slots = conv_tc_stat_slots(op.tc); f.fin_in = fr->mode == 2 ? fr->gn.stats : nullptr;
      f.fin_out = fr->stats_out; f.fin_HW = out.H * out.W; f.fin_C = a.Cout_p; f.fin_mode = fr->mode;
      f.bytes = 8.0 * B * f.fin_slots * 32; f.dims[0] = out.H; f.dims[1] = out.W; f.dims[2] = a.Cout_p;
      ops.push_back(f);
    }
    return fused;
  }
  double flops_acc = 0;
  // OSM_GN_COEF_BATCH (default 1): the backward-statistics coefficient sets in one launch per input-VJP, and the forward coefficients
  // written by the finalize launch that produces their statistics (together ~110 fewer launches per step)
  int use_coef_batch = [] { const char* e = getenv("OSM_GN_COEF_BATCH"); return e ? atoi(e) : 1; }();
  static constexpr int COEF_TABLE_CAP = 256;
  static constexpr size_t SPLITK_WS_FLOATS = (size_t)16 << 20;   // 64 MB: every split plan of the shipped configs (see conv_tc.cu)
  static constexpr int GN_GROUPS_ = 32;
  int use_stat_combine = [] { const char* e = getenv("OSM_GN_COMBINE"); return e ? atoi(e) : 1; }();

  // GroupNorm + SiLU fused into the operand load of the 3x3 conv that follows (halo kernel).  OSM_GN_XFORM: 0 off,
  // 1 (default) where the halo kernel fills the GPU (>= halo_min_tiles CTA-pair tiles), 2 wherever its shapes allow.
  int use_xform = [] { const char* e = getenv("OSM_GN_XFORM"); return e ? atoi(e) : 1; }();
  // Measured on B200 (tools/halo_probe.py): the halo kernel beats the tap-shifted pair kernel by 1.5-5 % once there is more
  // than one wave of pair tiles, and loses ~7 % on a single partial wave (64 tiles at 128x128, B = 1); with the operand
  // transform it costs +8 % on the conv and saves the whole GroupNorm apply pass (net -14 us per conv at 256x256, B = 1).
  int halo_min_tiles = [] { const char* e = getenv("OSM_HALO_MIN_TILES"); return e ? atoi(e) : 74; }();
  int use_halo = [] { const char* e = getenv("OSM_CONV_HALO"); return e ? atoi(e) : 1; }();   // plain halo kernel for the other 3x3 convs
  bool halo_wanted(int Hh, int Ww, int Cin_p, int Cout_p, int taps) const {
    if (conv_mode != 0 || !use_halo || !conv_tc_halo_ok(B, Hh, Ww, Cin_p, Cout_p, taps)) return false;
    const long ptiles = (((long)(Ww / 8) * (Hh / 16) * B + 1) / 2) * (Cout_p / 256);
    return use_halo == 2 || ptiles >= halo_min_tiles;
  }
  // fp16-operand halo kernel (OSM_CONV_F16: 0 off, 1 (default) from f16_min_tiles CTA-pair tiles, 2 wherever the shapes allow).
  // TF32 and fp16 have the same significand; the kernel runs at twice the tensor rate for the same shared-memory traffic.
  int use_f16 = [] { const char* e = getenv("OSM_CONV_F16"); return e ? atoi(e) : 1; }();
  // Below 17 pair tiles (batch 1: the 32x32 level with 1024 channels = 16 pairs on 32 of the 148 SMs, each walking the whole K loop:
  // 46-72 us per launch) the cluster split-K kernel on an fp16 operand written by the one-launch GroupNorm is ~2x faster since its
  // reduction goes through L2 (batch 1: 12.44 -> 12.22 ms per step; 33 measured the same as 17: the 64x64 level is a wash).
  int f16_min_tiles = [] { const char* e = getenv("OSM_F16_MIN_TILES"); return e ? atoi(e) : 17; }();
  bool f16_wanted(int Hh, int Ww, int Cin_p, int Cout_p, int taps) const {
    if (conv_mode != 0 || !use_f16 || !conv_tc_halo16_ok(B, Hh, Ww, Cin_p, Cout_p, taps)) return false;
    const long ptiles = (((long)(Ww / 8) * (Hh / 16) * B + 1) / 2) * (Cout_p == 64 ? 1 : Cout_p / 256);
    return use_f16 == 2 || ptiles >= f16_min_tiles;
  }
  // fp16 operands for the convs the halo kernels do not take (small images, 1x1 convs): possible when the conv's input has ONE
  // producer of ours that can write it as fp16 (GroupNorm apply / GroupNorm backward); decided before that producer is emitted.
  int use_nh16 = [] { const char* e = getenv("OSM_CONV_NH16"); return e ? atoi(e) : 1; }();
  bool nh16(int conv_idx, bool dgrad, int Hh, int Ww) const {
    if (conv_mode != 0 || !use_f16 || !use_nh16) return false;
    const ConvLayer& cl = convs[conv_idx];
    const int Cin_p = dgrad ? cl.Cout_p : cl.Cin_p, Cout_p = dgrad ? cl.Cin_p : cl.Cout_p;
    if (!(dgrad ? cl.wd16 : cl.wf16) || Cin_p % 64 != 0) return false;
    if (cl.taps == 9 && (f16_wanted(Hh, Ww, Cin_p, Cout_p, 9) || halo_wanted(Hh, Ww, Cin_p, Cout_p, 9))) return false;
    return true;
  }
  bool xform_wanted(int Hh, int Ww, int cin, int cout) const {
    if (conv_mode != 0 || !use_xform) return false;
    if (cin % 64 == 0 && cout % 64 == 0 && f16_wanted(Hh, Ww, cin, cout, 9)) return true;   // the fp16 kernel always transforms
    const int Cin_p = pad32(cin), Cout_p = pad32(cout);
    if (Cin_p != cin || Cin_p > 1536 || !conv_tc_halo_ok(B, Hh, Ww, Cin_p, Cout_p, 9)) return false;
    const long ptiles = (((long)(Ww / 8) * (Hh / 16) * B + 1) / 2) * (Cout_p / 256);
    return use_xform == 2 || ptiles >= halo_min_tiles;
  }
  // statistics (stand-alone pass unless the producer's epilogue reduced them) + the per-(image, channel) coefficients
  float* emit_gn_coef_fwd(PlanCtx& c, std::vector<Op>& ops, const GnArgs& a, bool have_stats) {
    float* coef = c.ar.alloc((size_t)B * a.C * 2);
    if (!have_stats) {
      const double n = (double)B * a.H * a.W * a.C;
      Op s{}; s.kind = OP_GN_STATS; s.gn = a; s.bytes = 4.0 * n; s.dims[0] = a.H; s.dims[1] = a.W; s.dims[2] = a.C; ops.push_back(s);
    }
    // the statistics were just finalized from a conv epilogue's partials: that launch writes the coefficients as well
    if (have_stats && use_coef_batch && !c.dry && !ops.empty() && ops.back().kind == OP_GN_FINALIZE && ops.back().fin_mode == 1 &&
        ops.back().fin_out == a.stats && ops.back().fin_C == a.C && !ops.back().fin_has_coef) {
      ops.back().fin_has_coef = true; ops.back().fin_gn = a; ops.back().fin_coef = coef;
      return coef;
    }
    Op k{}; k.kind = OP_GN_COEF_FWD; k.gn = a; k.gn_coef = coef; k.bytes = 8.0 * B * a.C;
    k.dims[0] = a.H; k.dims[1] = a.W; k.dims[2] = a.C;
    ops.push_back(k);
    return coef;
  }

  GnArgs make_gn(PlanCtx& c, View x, int g, int b, const float* ss, int silu, int resample, float* stats) {
    GnArgs a{};
    a.x = x.p; a.ldx = x.ld; a.gamma = raws[g].dev; a.beta = raws[b].dev;
    a.scale_shift = ss; a.ld_ss = emb_total; a.silu = silu; a.resample = resample; a.stats = stats;
    a.partial = c.partial; a.counter = c.counter; a.B = B; a.H = x.H; a.W = x.W; a.C = x.C;
    a.round_tf32 = conv_mode == 0;
    return a;
  }
  void emit_gn_fwd(PlanCtx& c, std::vector<Op>& ops, const GnArgs& a, float* y, bool have_stats = false) {
    const double n = (double)B * a.H * a.W * a.C;
    const double no = a.resample == RS_DOWN ? n / 4 : (a.resample == RS_UP ? n * 4 : n);
    static const int small_on = [] { const char* e = getenv("OSM_GN_SMALL"); return e ? atoi(e) : 1; }();
    const bool small = !have_stats && small_on && gn_small_capable(a);
    if (!have_stats && !small) {
      Op s{}; s.kind = OP_GN_STATS; s.gn = a; s.bytes = 4.0 * n; s.dims[0] = a.H; s.dims[1] = a.W; s.dims[2] = a.C; ops.push_back(s);
    }
    Op p{}; p.kind = OP_GN_APPLY; p.gn = a; p.gn_y = y; p.bytes = 4.0 * (n + no); p.dims[0] = a.H; p.dims[1] = a.W; p.dims[2] = a.C;
    p.gn_small = small;
    ops.push_back(p);
  }
  void emit_gn_bwd(PlanCtx& c, std::vector<Op>& ops, const GnArgs& f, const float* dy, View addend, int add_mode, View dx, int acc,
                   bool apply_only = false, bool dx_f16 = false) {
    static const int small_on = [] { const char* e = getenv("OSM_GN_SMALL"); return e ? atoi(e) : 1; }();
    Op o{}; o.kind = OP_GN_BWD; o.gnb_apply_only = apply_only;
    o.gn_small = !apply_only && small_on && gn_small_capable(f);
    o.gnb.f = f; o.gnb.dy = dy; o.gnb.addend = addend.p; o.gnb.ld_add = addend.ld; o.gnb.add_mode = add_mode;
    o.gnb.dx = dx.p; o.gnb.ld_dx = dx.ld; o.gnb.accumulate = acc; o.gnb.bstats = c.bstats; o.gnb.dx_f16 = dx_f16;
    {
      const double n = (double)B * f.H * f.W * f.C;
      const double ndy = f.resample == RS_DOWN ? n / 4 : (f.resample == RS_UP ? n * 4 : n);
      // two passes over (x, dy) (one when the means come from the conv epilogue) + write dx (+ addend / accumulate reads)
      o.bytes = 4.0 * ((apply_only ? 1 : 2) * (n + ndy) + n * (1 + (add_mode != ADD_NONE) + (acc != 0)));
      o.dims[0] = f.H; o.dims[1] = f.W; o.dims[2] = f.C;
    }
    ops.push_back(o);
  }

  static View dense(PlanCtx& c, int B, int H, int W, int C) {
    View v; v.p = c.ar.alloc((size_t)B * H * W * C); v.C = C; v.ld = C; v.H = H; v.W = W;
    return v;
  }
  static void need(size_t& slot, size_t n) { if (n > slot) slot = n; }

  // forward ops of one layer; records what its backward needs.  x/gx: input and its gradient view; y/gy: output views.
  void plan_layer(PlanCtx& c, const Layer& l, View x, View gx, View y, View gy) {
    std::vector<Op>& fw = fwd;
    const size_t px = (size_t)B * x.H * x.W, py = (size_t)B * y.H * y.W;
    PlanCtx::Rec rec{};
    rec.l = &l; rec.x = x; rec.gx = gx; rec.y = y; rec.gy = gy;
    // Statistics of the layer output y, reduced in the epilogue of the conv that produces it: the GroupNorm of the next
    // layer (same view, same channel count) then skips its own statistics pass.  (A skip-concat input is a different
    // (pointer, channels) key and keeps the stand-alone kernel.)
    float* st_y = c.ar.alloc((size_t)B * 64);
    FuseReq fy; fy.mode = 1; fy.stats_out = st_y;
    // A skip-concat input [h | hs] with equal halves whose producers both reduced their statistics: combine them (32 threads)
    // instead of a pass over the tensor.  (Unequal halves do not align with the concat's groups and keep the pass.)
    float* comb = c.ar.alloc((size_t)B * 64);   // allocated in the sizing passes too (they cannot see the pointers-keyed map)
    auto input_stats = [&](float* own) -> float* {
      auto it = c.fused_stats.find({x.p, x.C});
      if (it != c.fused_stats.end()) return it->second;
      if (use_stat_combine && x.p && x.C % 2 == 0 && (x.C / 2) % GN_GROUPS_ == 0) {
        auto ia = c.fused_stats.find({x.p, x.C / 2});
        auto ib = c.fused_stats.find({x.p + x.C / 2, x.C / 2});
        if (ia != c.fused_stats.end() && ib != c.fused_stats.end()) {
          Op k{}; k.kind = OP_GN_COMBINE; k.cmb_a = ia->second; k.cmb_b = ib->second; k.cmb_out = comb;
          k.bytes = 3.0 * B * 256; k.dims[0] = x.H; k.dims[1] = x.W; k.dims[2] = x.C;
          fw.push_back(k);
          c.fused_stats[{x.p, x.C}] = comb;
          return comb;
        }
      }
      return own;
    };
    bool y_fused = false;
    if (l.kind == L_CONV_IN) {
      View xin_v{xin, io_pad(), io_pad(), x.H, x.W};
      y_fused = emit_conv(c, fw, l.conv1, false, xin_v, y, convs[l.conv1].bias, View{}, RES_NONE, 0, &fy);
      need(c.need.sa, px * (size_t)io_pad());
    } else if (l.kind == L_RES) {
      const float* ss = c.embout ? c.embout + l.emb_off : nullptr;
      float* st1 = c.ar.alloc((size_t)B * 64);
      float* st2 = c.ar.alloc((size_t)B * 64);
      View h1 = dense(c, B, y.H, y.W, l.cout);
      need(c.need.sa, py * (size_t)std::max(l.cin, l.cout));
      need(c.need.sb, py * (size_t)l.cout);
      float* s1 = input_stats(st1);
      GnArgs gn1 = make_gn(c, x, l.g1, l.b1, nullptr, 1, l.updown, s1);
      FuseReq f2; f2.mode = 1; f2.stats_out = st2;
      const bool xf1 = l.updown == RS_NONE && xform_wanted(y.H, y.W, l.cin, l.cout);
      const bool xf2 = xform_wanted(y.H, y.W, l.cout, l.cout);
      bool h1_fused;
      if (xf1) {   // conv1 reads the raw block input: SiLU(GN(x)) happens in its operand load
        const float* cf1 = emit_gn_coef_fwd(c, fw, gn1, s1 != st1);
        h1_fused = emit_conv(c, fw, l.conv1, false, x, h1, convs[l.conv1].bias, View{}, RES_NONE, 0, &f2, cf1, 1);
      } else {
        const bool x16 = nh16(l.conv1, false, y.H, y.W);   // the GroupNorm pass writes the conv operand as fp16
        GnArgs g1 = gn1; g1.out_f16 = x16;
        emit_gn_fwd(c, fw, g1, c.SA, s1 != st1);
        View a1{c.SA, l.cin, l.cin, y.H, y.W};
        h1_fused = emit_conv(c, fw, l.conv1, false, a1, h1, convs[l.conv1].bias, View{}, RES_NONE, 0, &f2, nullptr, 0, false, x16);
      }
      GnArgs gn2 = make_gn(c, h1, l.g2, l.b2, ss, 1, RS_NONE, st2);
      View a2{c.SA, l.cout, l.cout, y.H, y.W};
      const float* cf2 = nullptr;
      bool x16_2 = false;
      if (xf2) { cf2 = emit_gn_coef_fwd(c, fw, gn2, h1_fused); a2 = h1; }   // conv2 reads the raw h1
      else {
        x16_2 = nh16(l.conv2, false, y.H, y.W);
        GnArgs g2 = gn2; g2.out_f16 = x16_2;
        emit_gn_fwd(c, fw, g2, c.SA, h1_fused);
      }
      if (l.skip >= 0) {
        emit_conv(c, fw, l.skip, false, x, y, convs[l.skip].bias, View{}, RES_NONE, 0);
        y_fused = emit_conv(c, fw, l.conv2, false, a2, y, convs[l.conv2].bias, y, RES_SAME, 0, &fy, cf2, 1, false, x16_2);
      } else {
        const int rm = l.updown == RS_DOWN ? RES_AVGPOOL : (l.updown == RS_UP ? RES_NEAREST_UP : RES_SAME);
        y_fused = emit_conv(c, fw, l.conv2, false, a2, y, convs[l.conv2].bias, x, rm, 0, &fy, cf2, 1, false, x16_2);
      }
      rec.gn1 = gn1; rec.gn2 = gn2; rec.h1 = h1;
    } else {  // attention
      const int L = x.H * x.W, C = l.cin;
      float* st = c.ar.alloc((size_t)B * 64);
      View qkv = dense(c, B, x.H, x.W, 3 * C);
      need(c.need.sa, px * (size_t)C);
      need(c.need.sb, px * (size_t)3 * C);
      // Product mode: fused tcgen05 kernels (no L x L matrix in memory).  Exact mode and shapes the fused kernels do not
      // take (channels per head != 64, L % 64 != 0) use the fp32-accurate batched-GEMM path with P / D scratch.
      // (the 8x8 level, L = 64: one fused fp32 launch per direction beats the 2 + 3 launches of the tcgen05 path - attention.cu)
      const bool flash = conv_mode == 0 && use_flash_attention && attn_flash_supported(L, C, l.heads) && !attention_small_ok(L, C, l.heads);
      View ao{c.SA, C, C, x.H, x.W};
      float *qkvT = nullptr, *lse = nullptr, *Dv = nullptr;
      if (flash) {
        qkvT = c.ar.alloc(px * (size_t)3 * C);
        ao = dense(c, B, x.H, x.W, C);
        lse = c.ar.alloc((size_t)B * l.heads * L);
        Dv = c.ar.alloc((size_t)B * l.heads * L);
        need(c.need.st, px * (size_t)C);
      } else {
        need(c.need.pd, (size_t)B * l.heads * L * L);
      }
      float* s_in = input_stats(st);
      GnArgs gn = make_gn(c, x, l.g1, l.b1, nullptr, 0, RS_NONE, s_in);
      const bool x16 = nh16(l.qkv, false, x.H, x.W);
      { GnArgs g = gn; g.out_f16 = x16; emit_gn_fwd(c, fw, g, c.SA, s_in != st); }
      View n{c.SA, C, C, x.H, x.W};
      emit_conv(c, fw, l.qkv, false, n, qkv, convs[l.qkv].bias, View{}, RES_NONE, 0, nullptr, nullptr, 0, false, x16);
      {
        Op o{}; o.kind = OP_ATTN_FWD; o.at_qkv = qkv.p; o.at_out = ao.p; o.at_L = L; o.at_C = C; o.at_heads = l.heads;
        o.flops = 4.0 * B * (double)L * L * C; o.bytes = 4.0 * B * (double)L * 4 * C; o.dims[0] = L; o.dims[1] = C; o.dims[2] = l.heads;
        o.at_flash = flash;
        if (flash && !c.dry) {
          if (int e = attn_flash_plan(&o.fa, qkv.p, qkvT, ao.p, lse, Dv, c.SA, c.ST, c.SB, B, L, C, l.heads)) c.err = e;
          rec.fa = o.fa;
        }
        fw.push_back(o);
        flops_acc += 4.0 * B * (double)L * L * C;
      }
      y_fused = emit_conv(c, fw, l.proj, false, ao, y, convs[l.proj].bias, x, RES_SAME, 0, &fy);
      rec.gn1 = gn; rec.qkv = qkv; rec.flash = flash;
    }
    if (y_fused) c.fused_stats[{y.p, y.C}] = st_y;
    c.recs.push_back(rec);
  }

  // backward ops of one recorded layer, called in reverse layer order so that `gwritten` reflects execution order
  void plan_layer_bwd(PlanCtx& c, const PlanCtx::Rec& r) {
    const Layer& l = *r.l;
    std::vector<Op>& bw = bwd;
    const View &x = r.x, &gx = r.gx, &y = r.y, &gy = r.gy;
    if (l.kind == L_CONV_IN) {
      View gxin_v{c.SA, io_pad(), io_pad(), x.H, x.W};
      emit_conv(c, bw, l.conv1, true, gy, gxin_v, nullptr, View{}, RES_NONE, 0);
      return;
    }
    bool& written = c.gwritten[gx.p];
    if (l.kind == L_RES) {
      // The two means every GroupNorm backward needs are reduced in the epilogue of the dgrad conv that produces its
      // dy (mode 2), so the GroupNorm backward is a single pass; resampling GroupNorms keep the two-kernel form.
      View t0{c.SA, l.cout, l.cout, y.H, y.W};
      FuseReq b2; b2.mode = 2; b2.stats_out = c.bstats; b2.gn = r.gn2;
      const bool f2 = emit_conv(c, bw, l.conv2, true, gy, t0, nullptr, View{}, RES_NONE, 0, &b2);
      View t1{c.SB, l.cout, l.cout, y.H, y.W};
      const bool t1_16 = nh16(l.conv1, true, y.H, y.W);    // the GroupNorm backward writes the dgrad conv's operand as fp16
      emit_gn_bwd(c, bw, r.gn2, c.SA, View{}, ADD_NONE, t1, 0, f2, t1_16);
      View t2{c.SA, l.cin, l.cin, y.H, y.W};
      FuseReq b1; b1.mode = 2; b1.stats_out = c.bstats; b1.gn = r.gn1;
      const bool f1 = emit_conv(c, bw, l.conv1, true, t1, t2, nullptr, View{}, RES_NONE, 0, &b1, nullptr, 0, false, t1_16);
      if (l.skip >= 0) {
        emit_conv(c, bw, l.skip, true, gy, gx, nullptr, View{}, RES_NONE, written ? 1 : 0);
        emit_gn_bwd(c, bw, r.gn1, c.SA, View{}, ADD_NONE, gx, 1, f1);
      } else {
        const int am = l.updown == RS_DOWN ? ADD_FROM_COARSE_QUARTER : (l.updown == RS_UP ? ADD_SUM4_FINE : ADD_SAME);
        emit_gn_bwd(c, bw, r.gn1, c.SA, gy, am, gx, written ? 1 : 0, f1);
      }
    } else {
      const int L = x.H * x.W, C = l.cin;
      View ga{c.SA, C, C, x.H, x.W};
      emit_conv(c, bw, l.proj, true, gy, ga, nullptr, View{}, RES_NONE, 0);
      {
        Op o{}; o.kind = OP_ATTN_BWD; o.at_qkv = r.qkv.p; o.at_g = c.SA; o.at_out = c.SB; o.at_L = L; o.at_C = C; o.at_heads = l.heads;
        o.flops = 10.0 * B * (double)L * L * C; o.bytes = 4.0 * B * (double)L * 8 * C; o.dims[0] = L; o.dims[1] = C; o.dims[2] = l.heads;
        o.at_flash = r.flash; o.fa = r.fa;
        bw.push_back(o);
      }
      View gq{c.SB, 3 * C, 3 * C, x.H, x.W};
      FuseReq bq; bq.mode = 2; bq.stats_out = c.bstats; bq.gn = r.gn1;
      const bool fq = emit_conv(c, bw, l.qkv, true, gq, ga, nullptr, View{}, RES_NONE, 0, &bq);
      emit_gn_bwd(c, bw, r.gn1, c.SA, gy, ADD_SAME, gx, written ? 1 : 0, fq);
    }
    written = true;
    auto it = c.cat_partner.find(gx.p);
    if (it != c.cat_partner.end() && gx.C == gx.ld) c.gwritten[it->second] = true;  // wrote the whole concat gradient
  }

  // Runs the planning pass.  dry: only sizes (no device pointers dereferenced, no tensor maps).
  int plan(int B_, int H_, int W_, void* ws, size_t* bytes_out, bool dry, const Scratch* sizes) {
    B = B_; H = H_; W = W_;
    const int mc = cfg.model_channels, ted = 4 * mc;
    const int down = 1 << (cfg.num_levels - 1);
    if (H % down || W % down) return fail(OSM_ERR_INVALID, "H and W must be divisible by 2^(levels-1)");
    fwd.clear(); bwd.clear(); flops_acc = 0;
    PlanCtx c{};
    c.e = this; c.ar.base = (char*)ws; c.dry = dry;
    // fixed buffers
    e0 = c.ar.alloc((size_t)B * mc);
    float* e1 = c.ar.alloc((size_t)B * ted);
    float* emb = c.ar.alloc((size_t)B * ted);
    c.embout = c.ar.alloc((size_t)B * emb_total);
    xin = c.ar.alloc((size_t)B * H * W * io_pad());
    yout = c.ar.alloc((size_t)B * H * W * io_pad());
    c.bstats = c.ar.alloc((size_t)B * 64);
    c.partial = (double*)c.ar.alloc((size_t)B * 1024 * 64 * 2);  // [B][<=1024 chunks][32 groups][2] doubles
    c.counter = (unsigned int*)c.ar.alloc((size_t)B + 64);
    vjp_amax = (unsigned int*)c.ar.alloc((size_t)B + 64);
    GnCoefDesc* coef_table = (GnCoefDesc*)c.ar.alloc((size_t)COEF_TABLE_CAP * sizeof(GnCoefDesc) / sizeof(float));
    if (conv_mode == 0) c.SK = c.ar.alloc(SPLITK_WS_FLOATS);
    if (sizes) {
      c.SA = c.ar.alloc(sizes->sa); c.SB = c.ar.alloc(sizes->sb); c.P = c.ar.alloc(sizes->pd); c.D = c.ar.alloc(sizes->pd);
      c.ST = c.ar.alloc(sizes->st);
      c.SP = c.ar.alloc(sizes->sp); c.SC = c.ar.alloc(sizes->sc);
      c.need_sp_alloc = sizes->sp;
    }
    gy = c.SB; gxin = c.SA;
    if (!dry) {
      OSM_CUDA_CHECK(cudaMemset(c.counter, 0, ((size_t)B + 64) * 4));
      OSM_CUDA_CHECK(cudaMemset(xin, 0, (size_t)B * H * W * io_pad() * 4));
    }
    // timestep MLP (unet.py:549-554) + all 42 emb_layers as ONE packed linear (unet.py:278-284)
    auto lin = [&](const float* in, int ldin, const float* w, const float* b, float* out, int ldout, int K, int N, int silu) {
      Op o{}; o.kind = OP_LINEAR; o.li_in = in; o.li_ldin = ldin; o.li_w = w; o.li_b = b; o.li_out = out; o.li_ldout = ldout;
      o.li_K = K; o.li_N = N; o.li_silu = silu;
      o.flops = 2.0 * B * K * N; o.bytes = 4.0 * ((double)K * N + (double)B * (K + N)); o.dims[0] = K; o.dims[1] = N;
      fwd.push_back(o);
      flops_acc += 2.0 * B * K * N;
    };
    lin(e0, mc, raws[te0_w].dev, raws[te0_b].dev, e1, ted, mc, ted, 0);
    lin(e1, ted, raws[te2_w].dev, raws[te2_b].dev, emb, ted, ted, ted, 1);
    lin(emb, ted, raws[emb_w_slot].dev, raws[emb_b_slot].dev, c.embout, emb_total, ted, emb_total, 1);

    // ---- shape inference for the skip-concat buffers ----
    struct Shp { int C, H, W; };
    std::vector<Shp> in_shape;
    {
      int h = H, w = W;
      for (auto& blk : in_blocks) {
        int ch = 0;
        for (int li : blk) {
          const Layer& l = layers[li];
          ch = l.cout;
          if (l.kind == L_RES && l.updown == RS_DOWN) { h /= 2; w /= 2; }
        }
        in_shape.push_back(Shp{ch, h, w});
      }
    }
    const int n_in = (int)in_blocks.size(), n_out = (int)out_blocks.size();
    if (n_in != n_out) return fail(OSM_ERR_STATE, "internal: block count mismatch");
    std::vector<View> cat(n_out), gcat(n_out), hs(n_in), ghs(n_in), hpart(n_out), ghpart(n_out);
    {
      int ch = in_shape.back().C;  // middle block keeps the channel count
      for (int j = 0; j < n_out; ++j) {
        const Shp s = in_shape[n_in - 1 - j];
        const int Ct = ch + s.C;
        cat[j] = dense(c, B, s.H, s.W, Ct);
        gcat[j] = dense(c, B, s.H, s.W, Ct);
        hpart[j] = View{cat[j].p, ch, Ct, s.H, s.W};
        ghpart[j] = View{gcat[j].p, ch, Ct, s.H, s.W};
        hs[n_in - 1 - j] = View{cat[j].p ? cat[j].p + ch : nullptr, s.C, Ct, s.H, s.W};
        ghs[n_in - 1 - j] = View{gcat[j].p ? gcat[j].p + ch : nullptr, s.C, Ct, s.H, s.W};
        if (gcat[j].p) c.cat_partner[gcat[j].p] = gcat[j].p + ch;
        ch = layers[out_blocks[j][0]].cout;
      }
    }

    auto run_block = [&](const std::vector<int>& blk, View x, View gx, View y_last, View gy_last) {
      for (size_t k = 0; k < blk.size(); ++k) {
        const Layer& l = layers[blk[k]];
        int oh = x.H, ow = x.W;
        if (l.kind == L_RES && l.updown == RS_DOWN) { oh /= 2; ow /= 2; }
        if (l.kind == L_RES && l.updown == RS_UP) { oh *= 2; ow *= 2; }
        View y, gyv;
        if (k + 1 == blk.size()) { y = y_last; gyv = gy_last; }
        else { y = dense(c, B, oh, ow, l.cout); gyv = dense(c, B, oh, ow, l.cout); }
        plan_layer(c, l, x, gx, y, gyv);
        x = y; gx = gyv;
      }
    };

    View x0v{nullptr, cfg.in_channels, io_pad(), H, W};
    run_block(in_blocks[0], x0v, View{}, hs[0], ghs[0]);
    for (int i = 1; i < n_in; ++i) run_block(in_blocks[i], hs[i - 1], ghs[i - 1], hs[i], ghs[i]);
    run_block(mid_block, hs[n_in - 1], ghs[n_in - 1], hpart[0], ghpart[0]);
    View hfinal = dense(c, B, H, W, layers[out_blocks[n_out - 1].back()].cout);
    View ghfinal = dense(c, B, H, W, hfinal.C);
    for (int j = 0; j < n_out; ++j) {
      View y = j + 1 < n_out ? hpart[j + 1] : hfinal;
      View gyv = j + 1 < n_out ? ghpart[j + 1] : ghfinal;
      run_block(out_blocks[j], cat[j], gcat[j], y, gyv);
    }
    // out: GroupNorm -> SiLU -> conv3x3 (unet.py:690-695, :742)
    {
      float* st = c.ar.alloc((size_t)B * 64);
      need(c.need.sa, (size_t)B * H * W * hfinal.C);
      need(c.need.sb, (size_t)B * H * W * io_pad());
      auto itf = c.fused_stats.find({hfinal.p, hfinal.C});
      float* s_out = itf != c.fused_stats.end() ? itf->second : st;
      GnArgs gn = make_gn(c, hfinal, out_g, out_b, nullptr, 1, RS_NONE, s_out);
      View yv{yout, io_pad(), io_pad(), H, W};
      if (use_xform && hfinal.C % 64 == 0 && f16_wanted(H, W, hfinal.C, convs[conv_out].Cout_p, 9)) {
        // the output conv reads the raw h and applies SiLU(GroupNorm(h)) in its operand load (fp16 halo kernel, 64-channel tile)
        const float* cfo = emit_gn_coef_fwd(c, fwd, gn, s_out != st);
        emit_conv(c, fwd, conv_out, false, hfinal, yv, convs[conv_out].bias, View{}, RES_NONE, 0, nullptr, cfo, 1);
      } else {
        const bool x16 = nh16(conv_out, false, H, W);
        { GnArgs g = gn; g.out_f16 = x16; emit_gn_fwd(c, fwd, g, c.SA, s_out != st); }
        View a{c.SA, hfinal.C, hfinal.C, H, W};
        emit_conv(c, fwd, conv_out, false, a, yv, convs[conv_out].bias, View{}, RES_NONE, 0, nullptr, nullptr, 0, false, x16);
      }
      // backward program: out layer first, then every layer in reverse
      View gyv{c.SB, io_pad(), io_pad(), H, W};
      View t0{c.SA, hfinal.C, hfinal.C, H, W};
      FuseReq bo; bo.mode = 2; bo.stats_out = c.bstats; bo.gn = gn;
      const bool fo = emit_conv(c, bwd, conv_out, true, gyv, t0, nullptr, View{}, RES_NONE, 0, &bo);
      emit_gn_bwd(c, bwd, gn, c.SA, View{}, ADD_NONE, ghfinal, 0, fo);
    }
    for (int g = (int)c.recs.size() - 1; g >= 0; --g) plan_layer_bwd(c, c.recs[g]);
    Pbuf = c.P; Dbuf = c.D;
    if (!dry && !c.coef_descs.empty()) {
      OSM_CUDA_CHECK(cudaMemcpy(coef_table, c.coef_descs.data(), c.coef_descs.size() * sizeof(GnCoefDesc), cudaMemcpyHostToDevice));
      Op k{}; k.kind = OP_GN_COEF_BATCH; k.cb_table = coef_table; k.cb_n = (int)c.coef_descs.size();
      k.bytes = 0; for (auto& d : c.coef_descs) k.bytes += 16.0 * B * d.C;
      bwd.insert(bwd.begin(), k);
    }

    if (c.err) return c.err;
    if (bytes_out) *bytes_out = c.ar.off + 256;
    last_need = c.need;
    flops = flops_acc;
    return OSM_OK;
  }
  Scratch last_need;

  int64_t workspace_bytes(int B_, int H_, int W_) {
    size_t bytes = 0;
    if (plan(B_, H_, W_, nullptr, &bytes, true, nullptr)) return -1;
    Scratch s = last_need;
    if (plan(B_, H_, W_, nullptr, &bytes, true, &s)) return -1;
    bound = false;
    return (int64_t)bytes;
  }

  int bind(int B_, int H_, int W_, void* ws, int64_t ws_bytes) {
    if (int e = ensure_storage()) return e;
    size_t bytes = 0;
    if (int e = plan(B_, H_, W_, nullptr, &bytes, true, nullptr)) return e;
    Scratch s = last_need;
    if (int e = plan(B_, H_, W_, nullptr, &bytes, true, &s)) return e;
    if ((int64_t)bytes > ws_bytes) return fail(OSM_ERR_INVALID, "workspace too small");
    if (((uintptr_t)ws) % 256) return fail(OSM_ERR_INVALID, "workspace must be 256-byte aligned");
    if (int e = plan(B_, H_, W_, ws, &bytes, false, &s)) return e;
    auto count = [&](const std::vector<Op>& ops) {
      int n = 0;
      for (auto& o : ops) {
        switch (o.kind) {
          case OP_GN_BWD: n += (o.gnb_apply_only || o.gn_small) ? 1 : 2; break;
          case OP_ATTN_FWD: n += o.at_flash ? 2 : (attention_small_ok(o.at_L, o.at_C, o.at_heads) ? 1 : attention_launches(0)); break;
          case OP_ATTN_BWD: n += o.at_flash ? 3 : (attention_small_ok(o.at_L, o.at_C, o.at_heads) ? 1 : attention_launches(1)); break;
          default: n += 1;
        }
      }
      return n;
    };
    fwd_launches = count(fwd) + 3;  // + layout-in, timestep embedding, layout-out
    bwd_launches = count(bwd) + 2 + (conv_mode == 0 && use_f16 ? 1 : 0);   // + amax of the cotangent
    bound = true;
    return OSM_OK;
  }

  int run(const Op& o, cudaStream_t s) {
    switch (o.kind) {
      case OP_CONV: return conv_mode == 0 ? conv_tc_launch(o.tc, s) : conv_simt_launch(o.tc.a, s);
      case OP_GN_STATS: return gn_stats_launch(o.gn, s);
      case OP_GN_APPLY: return o.gn_small ? gn_small_fwd_launch(o.gn, o.gn_y, s) : gn_apply_launch(o.gn, o.gn_y, s);
      case OP_GN_BWD:
        if (o.gn_small) return gn_small_bwd_launch(o.gnb, s);
        return o.gnb_apply_only ? gn_bwd_apply_launch(o.gnb, s) : gn_bwd_launch(o.gnb, s);
      case OP_GN_FINALIZE:
        return gn_fused_finalize_launch(o.fin_partial, o.fin_slots, o.fin_in, o.fin_out, B, o.fin_HW, o.fin_C, o.fin_mode, s,
                                        o.fin_has_coef ? &o.fin_gn : nullptr, o.fin_coef);
      case OP_GN_COEF_BATCH: return gn_coef_batch_launch(o.cb_table, o.cb_n, B, s);
      case OP_GN_COMBINE: return gn_combine_stats_launch(o.cmb_a, o.cmb_b, o.cmb_out, B, s);
      case OP_GN_COEF: return gn_coef_launch(o.gn, o.gn_coef, s);
      case OP_GN_COEF_FWD: return gn_coef_fwd_launch(o.gn, o.gn_coef, s);
      case OP_ATTN_FWD:
        return o.at_flash ? attn_flash_fwd_launch(o.fa, s) : attention_fwd_launch(o.at_qkv, o.at_out, Pbuf, B, o.at_L, o.at_C, o.at_heads, s);
      case OP_ATTN_BWD:
        return o.at_flash ? attn_flash_bwd_launch(o.fa, s)
                          : attention_bwd_launch(o.at_qkv, o.at_g, o.at_out, Pbuf, Dbuf, B, o.at_L, o.at_C, o.at_heads, s);
      case OP_LINEAR: return linear_launch(o.li_in, o.li_ldin, o.li_w, o.li_b, o.li_out, o.li_ldout, B, o.li_K, o.li_N, o.li_silu, s);
    }
    return fail(OSM_ERR_STATE, "unknown op");
  }
  float *Pbuf = nullptr, *Dbuf = nullptr;

  // Runs one program with a CUDA-event pair around every op (on the launching stream) and returns per-op device time.
  int profile(int which, cudaStream_t s, int cap, float* ms, int* kinds, double* fl, double* by, int* dims6) {
    if (!bound) return fail(OSM_ERR_STATE, "profile before bind");
    const std::vector<Op>& ops = which == 0 ? fwd : bwd;
    const int n = (int)ops.size();
    if (n > cap) return fail(OSM_ERR_INVALID, "profile: output arrays too small");
    std::vector<cudaEvent_t> ev(n + 1);
    for (auto& e : ev) OSM_CUDA_CHECK(cudaEventCreate(&e));
    OSM_CUDA_CHECK(cudaEventRecord(ev[0], s));
    for (int i = 0; i < n; ++i) {
      if (int e = run(ops[i], s)) return e;
      OSM_CUDA_CHECK(cudaEventRecord(ev[i + 1], s));
    }
    OSM_CUDA_CHECK(cudaStreamSynchronize(s));
    for (int i = 0; i < n; ++i) {
      OSM_CUDA_CHECK(cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]));
      kinds[i] = (int)ops[i].kind;
      fl[i] = ops[i].flops; by[i] = ops[i].bytes;
      for (int k = 0; k < 6; ++k) dims6[6 * i + k] = ops[i].dims[k];
    }
    for (auto& e : ev) cudaEventDestroy(e);
    return n;
  }

  int forward(const float* x, const float* t, float* out, cudaStream_t s) {
    if (!bound) return fail(OSM_ERR_STATE, "osm_unet_forward before osm_unet_bind");
    if (int e = nchw_to_nhwc_pad_launch(x, xin, B, cfg.in_channels, H * W, io_pad(), s)) return e;
    if (int e = timestep_embedding_launch(t, e0, B, cfg.model_channels, s)) return e;
    for (auto& o : fwd)
      if (int e = run(o, s)) return e;
    return nhwc_to_nchw_launch(yout, io_pad(), out, B, cfg.out_channels, H * W, s);
  }
  // The input-VJP is linear in grad_out.  With fp16-operand convs in the program every image's cotangent is multiplied by a power
  // of two on the way in (so that its largest entry lies in [2^vjp_texp, 2^(vjp_texp+1)) and the gradients inside the network sit in
  // the middle of the fp16 exponent range instead of its subnormal end: |c2 g| is ~1e-5 per pixel late in the chain) and the result
  // is divided by it on the way out; both are exact in fp32.
  unsigned int* vjp_amax = nullptr;
  int vjp_texp = [] { const char* e = getenv("OSM_VJP_SCALE_EXP"); return e ? atoi(e) : 4; }();
  int vjp(const float* grad_out, float* grad_x, cudaStream_t s) {
    if (!bound) return fail(OSM_ERR_STATE, "osm_unet_vjp_input before osm_unet_bind");
    const unsigned int* sb = nullptr;
    if (conv_mode == 0 && use_f16) {
      if (int e = amax_bits_launch(grad_out, vjp_amax, B, (size_t)cfg.out_channels * H * W, s)) return e;
      sb = vjp_amax;
    }
    if (int e = nchw_to_nhwc_pad_launch(grad_out, gy, B, cfg.out_channels, H * W, io_pad(), s, sb, vjp_texp)) return e;
    for (auto& o : bwd)
      if (int e = run(o, s)) return e;
    return nhwc_to_nchw_launch(gxin, io_pad(), grad_x, B, cfg.in_channels, H * W, s, sb, vjp_texp);
  }
};

}  // namespace osm

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
struct osm_unet {
  osm::Engine e;
};

extern "C" {

int osm_unet_create(const osm_unet_config* cfg, osm_unet_t* out) {
  if (!cfg || !out) return osm::fail(OSM_ERR_INVALID, "null argument");
  if (cfg->num_levels < 1 || cfg->num_levels > 8 || cfg->num_attention_ds < 0 || cfg->num_attention_ds > 8)
    return osm::fail(OSM_ERR_INVALID, "bad level / attention count");
  if (cfg->in_channels < 1 || cfg->in_channels > 32 || cfg->out_channels < 1 || cfg->out_channels > 32)
    return osm::fail(OSM_ERR_INVALID, "in/out channels must be in [1,32]");
  osm_unet* h = new osm_unet();
  h->e.cfg = *cfg;
  h->e.conv_mode = cfg->conv_mode;
  if (int e = h->e.build()) { delete h; return e; }
  *out = h;
  return OSM_OK;
}

int osm_unet_destroy(osm_unet_t h) {
  if (!h) return OSM_OK;
  if (h->e.wblock) cudaFree(h->e.wblock);
  if (h->e.stage) cudaFree(h->e.stage);
  delete h;
  return OSM_OK;
}

int osm_unet_param_count(osm_unet_t h) { return h ? (int)h->e.params.size() : -1; }

int osm_unet_param_info(osm_unet_t h, int index, const char** name, int* ndim, int64_t shape[4]) {
  if (!h || index < 0 || index >= (int)h->e.params.size()) return osm::fail(OSM_ERR_INVALID, "bad parameter index");
  const auto& p = h->e.params[index];
  if (name) *name = p.name.c_str();
  if (ndim) *ndim = (int)p.shape.size();
  if (shape)
    for (size_t i = 0; i < 4; ++i) shape[i] = i < p.shape.size() ? p.shape[i] : 1;
  return OSM_OK;
}

int osm_unet_load_param(osm_unet_t h, const char* name, const float* host_data, int64_t numel, void* stream) {
  if (!h || !name || !host_data) return osm::fail(OSM_ERR_INVALID, "null argument");
  return h->e.load_param(name, host_data, numel, (cudaStream_t)stream);
}

int64_t osm_unet_workspace_bytes(osm_unet_t h, int B, int H, int W) {
  if (!h || B < 1) return -1;
  return h->e.workspace_bytes(B, H, W);
}

int osm_unet_bind(osm_unet_t h, int B, int H, int W, void* workspace, int64_t workspace_bytes) {
  if (!h || !workspace || B < 1) return osm::fail(OSM_ERR_INVALID, "bad argument");
  return h->e.bind(B, H, W, workspace, workspace_bytes);
}

int osm_unet_forward(osm_unet_t h, const float* x, const float* t, float* out, void* stream) {
  if (!h) return osm::fail(OSM_ERR_INVALID, "null handle");
  return h->e.forward(x, t, out, (cudaStream_t)stream);
}

int osm_unet_vjp_input(osm_unet_t h, const float* grad_out, float* grad_x, void* stream) {
  if (!h) return osm::fail(OSM_ERR_INVALID, "null handle");
  return h->e.vjp(grad_out, grad_x, (cudaStream_t)stream);
}

int osm_unet_profile_ops(osm_unet_t h, int which, void* stream, int cap, float* ms, int* kinds, double* flops, double* bytes,
                         int* dims6) {
  if (!h || !ms || !kinds || !flops || !bytes || !dims6) return osm::fail(OSM_ERR_INVALID, "null argument");
  return h->e.profile(which, (cudaStream_t)stream, cap, ms, kinds, flops, bytes, dims6);
}

int osm_unet_launch_count(osm_unet_t h, int which) { return h ? (which == 0 ? h->e.fwd_launches : h->e.bwd_launches) : -1; }
double osm_unet_forward_flops(osm_unet_t h) { return h ? h->e.flops : 0.0; }

}  // extern "C"
