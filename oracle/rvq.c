/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see lyra_oracle.h).
 *
 * Residual vector quantizer restated from the two signatures of lyra/model_coeffs/quantizer.tflite
 * ("encode", "decode") that the reference drives through tflite::SignatureRunner
 * (lyra/residual_vector_quantizer.cc:45-61, 91-100, 129-166).  The graphs are 555 / 233 tiny ops;
 * instead of interpreting them op by op, the loader TRACES them and refuses to load unless they have
 * exactly the structure restated below, so the restatement cannot silently drift from the file:
 *
 *   encode, stage j (in residual-chain order, which must equal PACK order):
 *     d[c]  = SUM_j' SQUARED_DIFFERENCE(r, CB_j[c])      (f32; sum ascending over the 64 dims)
 *     i_j   = ARG_MIN_c d[c]                              (first minimum)
 *     q     = GATHER(CB_j, i_j);  t = q - r;  u = r + t;  r <- r - u      (three separate f32 ops)
 *     out_j = i_j if j < num_quantizers else -1           (ONE_HOT * mask, ARG_MAX, + (mask - 1))
 *   decode: out = (((t_0 + t_1) + t_2) + ...), t_k = GATHER(CB_k, max(idx_k, 0)) * (idx_k != -1)
 *
 * Assumed (not verifiable offline): ARG_MIN tie-break = lowest index; SUM over the
 * last axis accumulates in ascending order.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lyra_oracle.h"
#include "tfl_reader.h"

#define RVQ_MAX_STAGES 64
#define RVQ_DIM 64
#define RVQ_CODES 16

struct lo_rvq {
  int num_stages;
  int bits_per_stage;
  float cb[RVQ_MAX_STAGES][RVQ_CODES][RVQ_DIM];      /* encode-order codebooks */
  int dec_stage[RVQ_MAX_STAGES];                     /* decode ADD-chain position -> index slot */
  float dec_cb[RVQ_MAX_STAGES][RVQ_CODES][RVQ_DIM];  /* decode-chain codebooks */
};

static int* build_producers(const tfl_subgraph* sg) {
  int* prod = (int*)malloc(sizeof(int) * (size_t)sg->ntensors);
  for (int i = 0; i < sg->ntensors; ++i) prod[i] = -1;
  for (int o = 0; o < sg->nops; ++o)
    for (int k = 0; k < sg->ops[o].nout; ++k) prod[sg->ops[o].out[k]] = o;
  return prod;
}
static int through_reshape(const tfl_subgraph* sg, const int* prod, int t) {
  while (prod[t] >= 0 && sg->ops[prod[t]].code == OP_RESHAPE) t = sg->ops[prod[t]].in[0];
  return t;
}
static const tfl_op* producer(const tfl_subgraph* sg, const int* prod, int t, int code) {
  if (t < 0 || prod[t] < 0) return NULL;
  const tfl_op* op = &sg->ops[prod[t]];
  return op->code == code ? op : NULL;
}
/* STRIDED_SLICE picking one element along dim 0: returns begin[0], or -1 */
static int slice_begin0(const tfl_subgraph* sg, const tfl_op* op) {
  const tfl_tensor* b = &sg->tensors[op->in[1]];
  if (!b->data) return -1;
  return ((const int32_t*)b->data)[0];
}

#define FAIL(msg) do { fprintf(stderr, "lyra_oracle rvq: %s\n", msg); goto fail; } while (0)

lo_rvq* lo_rvq_create(const char* path) {
  tfl_model* m = tfl_load(path);
  if (!m) return NULL;
  lo_rvq* q = (lo_rvq*)calloc(1, sizeof(*q));
  int* prod = NULL;
  const int se = tfl_signature_subgraph(m, "encode"), sd = tfl_signature_subgraph(m, "decode");
  if (se < 0 || sd < 0) FAIL("missing encode/decode signature");

  /* ------------------------------------------------ encode ------------------------------- */
  {
    const tfl_subgraph* sg = &m->sub[se];
    prod = build_producers(sg);
    /* outputs: one i32[46,1,1] (indices), one i32 scalar constant (bits per quantizer) */
    int out_idx = -1;
    for (int i = 0; i < sg->nout; ++i) {
      const tfl_tensor* t = &sg->tensors[sg->outputs[i]];
      if (t->data && t->count == 1) q->bits_per_stage = *(const int32_t*)t->data;
      else out_idx = sg->outputs[i];
    }
    if (out_idx < 0 || q->bits_per_stage != 4) FAIL("unexpected encode outputs");
    int frames_in = -1;
    for (int i = 0; i < sg->nin; ++i)
      if (sg->tensors[sg->inputs[i]].type == TFL_F32) frames_in = sg->inputs[i];
    const tfl_op* add = producer(sg, prod, out_idx, OP_ADD);
    if (!add) FAIL("encode output is not ADD");
    const tfl_op* amax = producer(sg, prod, add->in[0], OP_ARG_MAX);
    if (!amax) FAIL("no ARG_MAX");
    const tfl_op* pack = producer(sg, prod, amax->in[0], OP_PACK);
    if (!pack || pack->nin > RVQ_MAX_STAGES) FAIL("no PACK");
    q->num_stages = pack->nin;
    int prev_res = -1, prev_argmin = -1, prev_cbg = -1;
    for (int j = 0; j < pack->nin; ++j) {
      const tfl_op* mul = producer(sg, prod, pack->in[j], OP_MUL);
      if (!mul) FAIL("PACK input is not MUL");
      const tfl_op* msl = producer(sg, prod, mul->in[1], OP_STRIDED_SLICE);
      if (!msl || slice_begin0(sg, msl) != j) FAIL("stage mask slice mismatch");
      const tfl_op* cast = producer(sg, prod, msl->in[0], OP_CAST);
      const tfl_op* less = cast ? producer(sg, prod, cast->in[0], OP_LESS) : NULL;
      if (!less || !sg->tensors[less->in[0]].data || ((const int32_t*)sg->tensors[less->in[0]].data)[j] != j)
        FAIL("stage mask is not (range < num_quantizers)");
      const tfl_op* oh = producer(sg, prod, through_reshape(sg, prod, mul->in[0]), OP_ONE_HOT);
      if (!oh) FAIL("no ONE_HOT");
      const tfl_op* amin = producer(sg, prod, oh->in[0], OP_ARG_MIN);
      if (!amin) FAIL("no ARG_MIN");
      const tfl_op* sum = producer(sg, prod, amin->in[0], OP_SUM);
      if (!sum) FAIL("no SUM");
      const tfl_op* sqd = producer(sg, prod, sum->in[0], OP_SQUARED_DIFFERENCE);
      if (!sqd) FAIL("no SQUARED_DIFFERENCE");
      const tfl_tensor* cbt = &sg->tensors[sqd->in[1]];
      if (!cbt->data || cbt->type != TFL_F32 || cbt->count != RVQ_CODES * RVQ_DIM) FAIL("codebook shape");
      memcpy(q->cb[j], cbt->data, sizeof(float) * RVQ_CODES * RVQ_DIM);
      const int res = sqd->in[0];
      if (j == 0) {
        if (through_reshape(sg, prod, res) != frames_in) FAIL("stage 0 residual is not the input");
      } else {
        /* r_j = SUB(r_{j-1}, ADD(r_{j-1}, SUB(RESHAPE(GATHER(CB_{j-1}, argmin_{j-1})), r_{j-1}))) */
        const tfl_op* s2 = producer(sg, prod, res, OP_SUB);
        if (!s2 || s2->in[0] != prev_res) FAIL("residual chain (outer SUB)");
        const tfl_op* a1 = producer(sg, prod, s2->in[1], OP_ADD);
        if (!a1 || a1->in[0] != prev_res) FAIL("residual chain (ADD)");
        const tfl_op* s1 = producer(sg, prod, a1->in[1], OP_SUB);
        if (!s1 || s1->in[1] != prev_res) FAIL("residual chain (inner SUB)");
        const tfl_op* ga = producer(sg, prod, through_reshape(sg, prod, s1->in[0]), OP_GATHER);
        if (!ga || ga->in[1] != prev_argmin) FAIL("residual chain (GATHER)");
        prev_cbg = ga->in[0];
        const tfl_tensor* g = &sg->tensors[prev_cbg];
        if (!g->data || memcmp(g->data, q->cb[j - 1], sizeof(float) * RVQ_CODES * RVQ_DIM) != 0)
          FAIL("GATHER codebook differs from SQUARED_DIFFERENCE codebook");
      }
      prev_res = res;
      prev_argmin = amin->out[0];
    }
    (void)prev_cbg;
    free(prod); prod = NULL;
  }
  /* ------------------------------------------------ decode ------------------------------- */
  {
    const tfl_subgraph* sg = &m->sub[sd];
    prod = build_producers(sg);
    int terms[RVQ_MAX_STAGES], nterms = 0;
    int cur = sg->outputs[0];
    const tfl_op* add;
    while ((add = producer(sg, prod, cur, OP_ADD)) != NULL) {
      if (nterms >= RVQ_MAX_STAGES - 1) FAIL("decode chain too long");
      terms[nterms++] = add->in[1];
      cur = add->in[0];
    }
    terms[nterms++] = cur;
    if (nterms != q->num_stages) FAIL("decode chain length != number of stages");
    for (int k = 0; k < nterms; ++k) {
      const int t = terms[nterms - 1 - k];   /* chain position k (leftmost first) */
      const tfl_op* mul = producer(sg, prod, t, OP_MUL);
      if (!mul) FAIL("decode term is not MUL");
      const tfl_op* ga = producer(sg, prod, through_reshape(sg, prod, mul->in[0]), OP_GATHER);
      if (!ga) FAIL("decode term has no GATHER");
      const tfl_tensor* g = &sg->tensors[ga->in[0]];
      if (!g->data || g->count != RVQ_CODES * RVQ_DIM) FAIL("decode codebook shape");
      memcpy(q->dec_cb[k], g->data, sizeof(float) * RVQ_CODES * RVQ_DIM);
      const tfl_op* isl = producer(sg, prod, ga->in[1], OP_STRIDED_SLICE);
      const tfl_op* msl = producer(sg, prod, mul->in[1], OP_STRIDED_SLICE);
      if (!isl || !msl) FAIL("decode slices");
      const int st = slice_begin0(sg, isl);
      if (st < 0 || st >= q->num_stages || slice_begin0(sg, msl) != st) FAIL("decode index/mask slot mismatch");
      if (!producer(sg, prod, isl->in[0], OP_MAXIMUM)) FAIL("decode index not MAXIMUM(idx, 0)");
      const tfl_op* cast = producer(sg, prod, msl->in[0], OP_CAST);
      if (!cast || !producer(sg, prod, cast->in[0], OP_NOT_EQUAL)) FAIL("decode mask not CAST(NOT_EQUAL)");
      q->dec_stage[k] = st;
    }
    free(prod); prod = NULL;
  }
  tfl_free(m);
  return q;
fail:
  free(prod);
  tfl_free(m);
  free(q);
  return NULL;
}

void lo_rvq_free(lo_rvq* q) { free(q); }
int lo_rvq_num_stages(const lo_rvq* q) { return q->num_stages; }
int lo_rvq_bits_per_stage(const lo_rvq* q) { return q->bits_per_stage; }
const float* lo_rvq_codebook(const lo_rvq* q, int stage) {
  return stage >= 0 && stage < q->num_stages ? &q->cb[stage][0][0] : NULL;
}

int lo_rvq_encode(const lo_rvq* q, const float* f, int num_quantizers, int32_t* indices) {
  float r[RVQ_DIM];
  memcpy(r, f, sizeof(r));
  /* the graph always evaluates all stages; stages >= num_quantizers report -1 */
  for (int s = 0; s < q->num_stages; ++s) {
    int best = 0;
    float bestd = 0.0f;
    for (int c = 0; c < RVQ_CODES; ++c) {
      float d = 0.0f;
      for (int j = 0; j < RVQ_DIM; ++j) {
        const float df = r[j] - q->cb[s][c][j];
        const float sq = df * df;
        d = d + sq;
      }
      if (c == 0 || d < bestd) { bestd = d; best = c; }
    }
    for (int j = 0; j < RVQ_DIM; ++j) {
      const float t = q->cb[s][best][j] - r[j];
      const float u = r[j] + t;
      r[j] = r[j] - u;
    }
    indices[s] = s < num_quantizers ? best : -1;
  }
  return 0;
}

int lo_rvq_decode(const lo_rvq* q, const int32_t* indices, float* out) {
  for (int k = 0; k < q->num_stages; ++k) {
    const int st = q->dec_stage[k];
    const int32_t idx = indices[st];
    if (idx > RVQ_CODES - 1) return -1;
    const float mask = idx != -1 ? 1.0f : 0.0f;
    const float* cb = q->dec_cb[k][idx > 0 ? idx : 0];
    for (int j = 0; j < RVQ_DIM; ++j) {
      const float t = cb[j] * mask;
      out[j] = k == 0 ? t : out[j] + t;
    }
  }
  return 0;
}

/* lyra/residual_vector_quantizer.h:50  kMaxNumQuantizedBits = 184 */
#define RVQ_MAX_BITS 184

int lo_rvq_quantize_bits(const lo_rvq* q, const float* f, int num_bits, char* bits_out) {
  if (num_bits > RVQ_MAX_BITS || num_bits < 0 || num_bits % q->bits_per_stage != 0) return -1;
  const int nq = num_bits / q->bits_per_stage;
  int32_t idx[RVQ_MAX_STAGES];
  lo_rvq_encode(q, f, nq, idx);
  /* first quantizer in the most significant bits (residual_vector_quantizer.cc:101-109) */
  for (int i = 0; i < nq; ++i)
    for (int b = 0; b < q->bits_per_stage; ++b)
      bits_out[i * q->bits_per_stage + b] = (idx[i] >> (q->bits_per_stage - 1 - b)) & 1 ? '1' : '0';
  bits_out[num_bits] = 0;
  return 0;
}

int lo_rvq_decode_bits(const lo_rvq* q, const char* bits, int num_bits, float* out) {
  if (num_bits > RVQ_MAX_BITS || num_bits < 0 || num_bits % q->bits_per_stage != 0) return -1;
  const int nq = num_bits / q->bits_per_stage;
  int32_t idx[RVQ_MAX_STAGES];
  for (int i = 0; i < nq; ++i) {
    int v = 0;
    for (int b = 0; b < q->bits_per_stage; ++b) v = (v << 1) | (bits[i * q->bits_per_stage + b] == '1');
    idx[i] = v;
  }
  /* unused quantizers are marked -1 (residual_vector_quantizer.cc:155-157) */
  for (int j = nq; j < q->num_stages; ++j) idx[j] = -1;
  return lo_rvq_decode(q, idx, out);
}
