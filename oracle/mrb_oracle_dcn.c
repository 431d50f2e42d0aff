/*
 * mrb_oracle_dcn.c -- CPU restatement of the reference's deformable conv (v1/v2) and deformable
 * PS-ROI pooling.  TEST INFRASTRUCTURE ONLY (see mrb_oracle.c header).
 *
 * The reference has no CPU implementation of these ops (csrc/deform_conv.h:41,76,111,148,190;
 * csrc/deform_pool.h:37,69) and no numeric test: "parity unpinned" by the reference's own tests.
 * The restatement serialises the CUDA kernels thread by thread in index order and replaces the
 * cuBLAS addmm calls of csrc/cuda/deform_conv_cuda.cu by plain triple loops; it is pinned in
 * tests/test_oracle.py against torchvision.ops.deform_conv2d (forward + autograd) on CPU.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

typedef struct {
  int batch, cin, H, W, cout, kh, kw, sh, sw, ph, pw, dh, dw, groups, dg;
} orc_dcn;

static int out_dim(int in, int pad, int dil, int k, int stride) { return (in + 2 * pad - (dil * (k - 1) + 1)) / stride + 1; }

/* deformable_im2col_bilinear == dmcn_im2col_bilinear, deform_conv_kernel_cuda.cu:91-121,473-503 */
static float im2col_bilinear(const float* d, int data_width, int height, int width, float h, float w) {
  int h_low = (int)floorf(h), w_low = (int)floorf(w);
  int h_high = h_low + 1, w_high = w_low + 1;
  float lh = h - h_low, lw = w - w_low, hh = 1 - lh, hw = 1 - lw;
  float v1 = 0, v2 = 0, v3 = 0, v4 = 0;
  if (h_low >= 0 && w_low >= 0) v1 = d[h_low * data_width + w_low];
  if (h_low >= 0 && w_high <= width - 1) v2 = d[h_low * data_width + w_high];
  if (h_high <= height - 1 && w_low >= 0) v3 = d[h_high * data_width + w_low];
  if (h_high <= height - 1 && w_high <= width - 1) v4 = d[h_high * data_width + w_high];
  float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
  return (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4);
}

/* get_gradient_weight == dmcn_get_gradient_weight, :123-147,506-531 */
static float gradient_weight(float ah, float aw, int h, int w, int height, int width) {
  if (ah <= -1 || ah >= height || aw <= -1 || aw >= width) return 0;
  int hl = (int)floorf(ah), wl = (int)floorf(aw), hh = hl + 1, wh = wl + 1;
  float weight = 0;
  if (h == hl && w == wl) weight = (h + 1 - ah) * (w + 1 - aw);
  if (h == hl && w == wh) weight = (h + 1 - ah) * (aw + 1 - w);
  if (h == hh && w == wl) weight = (ah + 1 - h) * (w + 1 - aw);
  if (h == hh && w == wh) weight = (ah + 1 - h) * (aw + 1 - w);
  return weight;
}

/* get_coordinate_weight == dmcn_get_coordinate_weight, :149-195,533-575 */
static float coordinate_weight(float ah, float aw, int height, int width, const float* im, int data_width, int bp_dir) {
  if (ah <= -1 || ah >= height || aw <= -1 || aw >= width) return 0;
  int hl = (int)floorf(ah), wl = (int)floorf(aw), hh = hl + 1, wh = wl + 1;
  float weight = 0;
  if (bp_dir == 0) {
    if (hl >= 0 && wl >= 0) weight += -1 * (wl + 1 - aw) * im[hl * data_width + wl];
    if (hl >= 0 && wh <= width - 1) weight += -1 * (aw - wl) * im[hl * data_width + wh];
    if (hh <= height - 1 && wl >= 0) weight += (wl + 1 - aw) * im[hh * data_width + wl];
    if (hh <= height - 1 && wh <= width - 1) weight += (aw - wl) * im[hh * data_width + wh];
  } else {
    if (hl >= 0 && wl >= 0) weight += -1 * (hl + 1 - ah) * im[hl * data_width + wl];
    if (hl >= 0 && wh <= width - 1) weight += (hl + 1 - ah) * im[hl * data_width + wh];
    if (hh <= height - 1 && wl >= 0) weight += -1 * (ah - hl) * im[hh * data_width + wl];
    if (hh <= height - 1 && wh <= width - 1) weight += (ah - hl) * im[hh * data_width + wh];
  }
  return weight;
}

/* (modulated_)deformable_im2col_gpu_kernel with batch_size == 1, :197-250,577-640.
 * im [C,H,W], offset [dg*2*taps,Ho,Wo], mask [dg*taps,Ho,Wo] or NULL, col [C*taps, Ho*Wo] */
static void im2col(const orc_dcn* p, const float* im, const float* offset, const float* mask, float* col) {
  const int Ho = out_dim(p->H, p->ph, p->dh, p->kh, p->sh), Wo = out_dim(p->W, p->pw, p->dw, p->kw, p->sw);
  const int taps = p->kh * p->kw, cpg = p->cin / p->dg;
  for (int c = 0; c < p->cin; c++) for (int h_col = 0; h_col < Ho; h_col++) for (int w_col = 0; w_col < Wo; w_col++) {
    const int d = c / cpg, h_in = h_col * p->sh - p->ph, w_in = w_col * p->sw - p->pw;
    const float* imc = im + (size_t)c * p->H * p->W;
    const float* off = offset + (size_t)d * 2 * taps * Ho * Wo;
    const float* msk = mask ? mask + (size_t)d * taps * Ho * Wo : NULL;
    for (int i = 0; i < p->kh; ++i) for (int j = 0; j < p->kw; ++j) {
      const int t = i * p->kw + j;
      const float oh = off[((2 * t) * Ho + h_col) * Wo + w_col], ow = off[((2 * t + 1) * Ho + h_col) * Wo + w_col];
      float val = 0;
      const float h_im = h_in + i * p->dh + oh, w_im = w_in + j * p->dw + ow;
      if (h_im > -1 && w_im > -1 && h_im < p->H && w_im < p->W) val = im2col_bilinear(imc, p->W, p->H, p->W, h_im, w_im);
      if (msk) val = val * msk[(t * Ho + h_col) * Wo + w_col];
      col[(((size_t)c * taps + t) * Ho + h_col) * Wo + w_col] = val;
    }
  }
}

/* (modulated_)deformable_col2im_gpu_kernel, :286-342,642-700: 5x5 window scan as in the reference */
static void col2im(const orc_dcn* p, const float* col, const float* offset, const float* mask, float* grad_im) {
  const int Ho = out_dim(p->H, p->ph, p->dh, p->kh, p->sh), Wo = out_dim(p->W, p->pw, p->dw, p->kw, p->sw);
  const int taps = p->kh * p->kw, cpg = p->cin / p->dg;
  for (int c = 0; c < p->cin; c++) for (int i = 0; i < p->kh; i++) for (int j = 0; j < p->kw; j++)
    for (int h_out = 0; h_out < Ho; h_out++) for (int w_out = 0; w_out < Wo; w_out++) {
      const int d = c / cpg, t = i * p->kw + j;
      const int w_in = w_out * p->sw - p->pw, h_in = h_out * p->sh - p->ph;
      const float* off = offset + (size_t)d * 2 * taps * Ho * Wo;
      const float oh = off[((2 * t) * Ho + h_out) * Wo + w_out], ow = off[((2 * t + 1) * Ho + h_out) * Wo + w_out];
      const float ch = h_in + i * p->dh + oh, cw = w_in + j * p->dw + ow;
      float top = col[(((size_t)c * taps + t) * Ho + h_out) * Wo + w_out];
      if (mask) top = top * mask[(((size_t)d * taps + t) * Ho + h_out) * Wo + w_out];
      const int cur_h = (int)ch, cur_w = (int)cw;
      for (int dy = -2; dy <= 2; dy++) for (int dx = -2; dx <= 2; dx++) {
        if (cur_h + dy >= 0 && cur_h + dy < p->H && cur_w + dx >= 0 && cur_w + dx < p->W &&
            fabsf(ch - (cur_h + dy)) < 1 && fabsf(cw - (cur_w + dx)) < 1) {
          float weight = gradient_weight(ch, cw, cur_h + dy, cur_w + dx, p->H, p->W);
          grad_im[((size_t)c * p->H + cur_h + dy) * p->W + cur_w + dx] += weight * top;
        }
      }
    }
}

/* (modulated_)deformable_col2im_coord_gpu_kernel, :380-443,702-774 */
static void col2im_coord(const orc_dcn* p, const float* col, const float* im, const float* offset, const float* mask,
                         float* grad_offset, float* grad_mask) {
  const int Ho = out_dim(p->H, p->ph, p->dh, p->kh, p->sh), Wo = out_dim(p->W, p->pw, p->dw, p->kw, p->sw);
  const int taps = p->kh * p->kw, cpg = p->cin / p->dg;
  for (int oc = 0; oc < 2 * taps * p->dg; oc++) for (int h = 0; h < Ho; h++) for (int w = 0; w < Wo; w++) {
    float val = 0, mval = 0;
    const int d = oc / (2 * taps), offset_c = oc - d * 2 * taps;
    const int t = offset_c / 2, bp_dir = offset_c % 2;
    const int i = t / p->kw, j = t % p->kw;
    const float* off = offset + (size_t)d * 2 * taps * Ho * Wo;
    for (int cc = 0; cc < cpg; cc++) {
      const int c = d * cpg + cc;
      const float cg = col[(((size_t)c * taps + t) * Ho + h) * Wo + w];
      const int w_in = w * p->sw - p->pw, h_in = h * p->sh - p->ph;
      const float oh = off[((2 * t) * Ho + h) * Wo + w], ow = off[((2 * t + 1) * Ho + h) * Wo + w];
      const float m = mask ? mask[(((size_t)d * taps + t) * Ho + h) * Wo + w] : 1.f;
      float inv_h = h_in + i * p->dh + oh, inv_w = w_in + j * p->dw + ow;
      const float* imc = im + (size_t)c * p->H * p->W;
      if (inv_h <= -1 || inv_w <= -1 || inv_h >= p->H || inv_w >= p->W) {
        inv_h = inv_w = -2;
      } else {
        mval += cg * im2col_bilinear(imc, p->W, p->H, p->W, inv_h, inv_w);
      }
      const float weight = coordinate_weight(inv_h, inv_w, p->H, p->W, imc, p->W, bp_dir);
      val += weight * cg * m;
    }
    grad_offset[((size_t)oc * Ho + h) * Wo + w] = val;
    if (grad_mask && offset_c % 2 == 0) grad_mask[(((size_t)d * taps + t) * Ho + h) * Wo + w] = mval;
  }
}

/* deform_conv_forward_cuda (deform_conv_cuda.cu:158-266) / modulated_deform_conv_cuda_forward
 * (:496-575): out[b] = W_g . columns_g (+ bias).  mask == NULL -> v1. */
ORC_API void orc_deform_conv_fwd(const orc_dcn* p, const float* input, const float* offset, const float* mask,
                                 const float* weight, const float* bias, float* output) {
  const int Ho = out_dim(p->H, p->ph, p->dh, p->kh, p->sh), Wo = out_dim(p->W, p->pw, p->dw, p->kw, p->sw);
  const int taps = p->kh * p->kw, P = Ho * Wo, Kg = p->cin / p->groups * taps, Mg = p->cout / p->groups;
  float* col = (float*)malloc(sizeof(float) * (size_t)p->cin * taps * P);
  for (int b = 0; b < p->batch; b++) {
    im2col(p, input + (size_t)b * p->cin * p->H * p->W, offset + (size_t)b * p->dg * 2 * taps * P,
           mask ? mask + (size_t)b * p->dg * taps * P : NULL, col);
    for (int g = 0; g < p->groups; g++) for (int m = 0; m < Mg; m++) {
      float* o = output + ((size_t)b * p->cout + g * Mg + m) * P;
      const float* wr = weight + ((size_t)g * Mg + m) * Kg;
      for (int x = 0; x < P; x++) o[x] = 0.f;
      for (int k = 0; k < Kg; k++) {
        const float wv = wr[k];
        const float* cr = col + ((size_t)g * Kg + k) * P;
        for (int x = 0; x < P; x++) o[x] += wv * cr[x];
      }
      if (bias) for (int x = 0; x < P; x++) o[x] += bias[g * Mg + m];
    }
  }
  free(col);
}

/* deform_conv_backward_input_cuda + deform_conv_backward_parameters_cuda (:268-494) /
 * modulated_deform_conv_cuda_backward (:577-691).  grad_input accumulates (caller zero-fills, as the
 * python wrappers do); grad_offset / grad_mask are overwritten; grad_weight += scale * dW;
 * grad_bias += sum.  Any output pointer may be NULL. */
ORC_API void orc_deform_conv_bwd(const orc_dcn* p, const float* input, const float* offset, const float* mask,
                                 const float* weight, const float* grad_output, float* grad_input,
                                 float* grad_offset, float* grad_mask, float* grad_weight, float* grad_bias,
                                 float scale) {
  const int Ho = out_dim(p->H, p->ph, p->dh, p->kh, p->sh), Wo = out_dim(p->W, p->pw, p->dw, p->kw, p->sw);
  const int taps = p->kh * p->kw, P = Ho * Wo, Kg = p->cin / p->groups * taps, Mg = p->cout / p->groups;
  float* col = (float*)malloc(sizeof(float) * (size_t)p->cin * taps * P);
  for (int b = 0; b < p->batch; b++) {
    const float* im = input + (size_t)b * p->cin * p->H * p->W;
    const float* off = offset + (size_t)b * p->dg * 2 * taps * P;
    const float* msk = mask ? mask + (size_t)b * p->dg * taps * P : NULL;
    const float* go = grad_output + (size_t)b * p->cout * P;
    if (grad_input || grad_offset) {
      for (int g = 0; g < p->groups; g++) for (int k = 0; k < Kg; k++) {
        float* cr = col + ((size_t)g * Kg + k) * P;
        for (int x = 0; x < P; x++) cr[x] = 0.f;
        for (int m = 0; m < Mg; m++) {
          const float wv = weight[((size_t)g * Mg + m) * Kg + k];
          const float* gr = go + ((size_t)g * Mg + m) * P;
          for (int x = 0; x < P; x++) cr[x] += wv * gr[x];
        }
      }
      if (grad_offset)
        col2im_coord(p, col, im, off, msk, grad_offset + (size_t)b * p->dg * 2 * taps * P,
                     grad_mask ? grad_mask + (size_t)b * p->dg * taps * P : NULL);
      if (grad_input) col2im(p, col, off, msk, grad_input + (size_t)b * p->cin * p->H * p->W);
    }
    if (grad_weight) {
      im2col(p, im, off, msk, col);
      for (int g = 0; g < p->groups; g++) for (int m = 0; m < Mg; m++) for (int k = 0; k < Kg; k++) {
        const float* gr = go + ((size_t)g * Mg + m) * P;
        const float* cr = col + ((size_t)g * Kg + k) * P;
        float s = 0.f;
        for (int x = 0; x < P; x++) s += gr[x] * cr[x];
        grad_weight[((size_t)g * Mg + m) * Kg + k] += scale * s;
      }
    }
    if (grad_bias) for (int m = 0; m < p->cout; m++) {
      float s = 0.f;
      for (int x = 0; x < P; x++) s += go[(size_t)m * P + x];
      grad_bias[m] += s;
    }
  }
  free(col);
}

/* ------------------------------------------------------- deformable PS-ROI pooling ---------- */
typedef struct {
  int channels, height, width, pooled, output_dim, group_size, part_size, sample_per_part, num_classes,
      channels_each_class, no_trans;
  float spatial_scale, trans_std;
} orc_ps;

/* deform_pool_kernel_cuda.cu:31-51 */
static float ps_bilinear(const float* data, float x, float y, int width, int height) {
  int x1 = (int)floorf(x), x2 = (int)ceilf(x), y1 = (int)floorf(y), y2 = (int)ceilf(y);
  float dx = x - x1, dy = y - y1;
  (void)height;
  return (1 - dx) * (1 - dy) * data[y1 * width + x1] + (1 - dx) * dy * data[y2 * width + x1] +
         dx * (1 - dy) * data[y1 * width + x2] + dx * dy * data[y2 * width + x2];
}

/* DeformablePSROIPoolForwardKernel / BackwardAccKernel, deform_pool_kernel_cuda.cu:53-264.
 * bwd != 0 runs the backward (accumulating into data_diff / trans_diff). */
static void ps_run(const orc_ps* a, int num_rois, const float* data, const float* rois, const float* trans,
                   float* top_data, float* top_count, int bwd, const float* top_diff, float* data_diff,
                   float* trans_diff) {
  const int count = num_rois * a->output_dim * a->pooled * a->pooled;
  for (int index = 0; index < count; index++) {
    int pw = index % a->pooled, ph = (index / a->pooled) % a->pooled;
    int ctop = (index / a->pooled / a->pooled) % a->output_dim, n = index / a->pooled / a->pooled / a->output_dim;
    const float* r = rois + n * 5;
    int bi = (int)r[0];
    float rsw = (float)(roundf(r[1]) * a->spatial_scale - 0.5), rsh = (float)(roundf(r[2]) * a->spatial_scale - 0.5);
    float rew = (float)((float)(roundf(r[3]) + 1.) * a->spatial_scale - 0.5);
    float reh = (float)((float)(roundf(r[4]) + 1.) * a->spatial_scale - 0.5);
    float roi_w = fmaxf(rew - rsw, 0.1f), roi_h = fmaxf(reh - rsh, 0.1f);
    float bin_h = roi_h / (float)a->pooled, bin_w = roi_w / (float)a->pooled;
    float sub_h = bin_h / (float)a->sample_per_part, sub_w = bin_w / (float)a->sample_per_part;
    int part_h = (int)floorf((float)ph / a->pooled * a->part_size), part_w = (int)floorf((float)pw / a->pooled * a->part_size);
    int class_id = ctop / a->channels_each_class;
    size_t tix = ((((size_t)n * a->num_classes + class_id) * 2) * a->part_size + part_h) * a->part_size + part_w;
    size_t tiy = ((((size_t)n * a->num_classes + class_id) * 2 + 1) * a->part_size + part_h) * a->part_size + part_w;
    float tx = a->no_trans ? 0.f : trans[tix] * a->trans_std, ty = a->no_trans ? 0.f : trans[tiy] * a->trans_std;
    float wstart = (float)pw * bin_w + rsw; wstart += tx * roi_w;
    float hstart = (float)ph * bin_h + rsh; hstart += ty * roi_h;
    int gw = (int)floorf((float)pw * a->group_size / a->pooled), gh = (int)floorf((float)ph * a->group_size / a->pooled);
    gw = gw < 0 ? 0 : (gw > a->group_size - 1 ? a->group_size - 1 : gw);
    gh = gh < 0 ? 0 : (gh > a->group_size - 1 ? a->group_size - 1 : gh);
    int c = (ctop * a->group_size + gh) * a->group_size + gw;
    const size_t base = ((size_t)bi * a->channels + c) * a->height * a->width;
    if (!bwd) {
      float sum = 0; int cnt = 0;
      for (int ih = 0; ih < a->sample_per_part; ih++) for (int iw = 0; iw < a->sample_per_part; iw++) {
        float w = wstart + iw * sub_w, h = hstart + ih * sub_h;
        if (w < -0.5 || w > a->width - 0.5 || h < -0.5 || h > a->height - 0.5) continue;
        w = fminf(fmaxf(w, 0.f), a->width - 1.f); h = fminf(fmaxf(h, 0.f), a->height - 1.f);
        sum += ps_bilinear(data + base, w, h, a->width, a->height); cnt++;
      }
      top_data[index] = cnt == 0 ? 0.f : sum / cnt;
      top_count[index] = (float)cnt;
    } else {
      if (top_count[index] <= 0) continue;
      float diff_val = top_diff[index] / top_count[index];
      for (int ih = 0; ih < a->sample_per_part; ih++) for (int iw = 0; iw < a->sample_per_part; iw++) {
        float w = wstart + iw * sub_w, h = hstart + ih * sub_h;
        if (w < -0.5 || w > a->width - 0.5 || h < -0.5 || h > a->height - 0.5) continue;
        w = fminf(fmaxf(w, 0.f), a->width - 1.f); h = fminf(fmaxf(h, 0.f), a->height - 1.f);
        int x0 = (int)floorf(w), x1 = (int)ceilf(w), y0 = (int)floorf(h), y1 = (int)ceilf(h);
        float dx = w - x0, dy = h - y0;
        data_diff[base + y0 * a->width + x0] += (1 - dx) * (1 - dy) * diff_val;
        data_diff[base + y1 * a->width + x0] += (1 - dx) * dy * diff_val;
        data_diff[base + y0 * a->width + x1] += dx * (1 - dy) * diff_val;
        data_diff[base + y1 * a->width + x1] += dx * dy * diff_val;
        if (a->no_trans) continue;
        float u00 = data[base + y0 * a->width + x0], u01 = data[base + y1 * a->width + x0];
        float u10 = data[base + y0 * a->width + x1], u11 = data[base + y1 * a->width + x1];
        float diff_x = (u11 * dy + u10 * (1 - dy) - u01 * dy - u00 * (1 - dy)) * a->trans_std * diff_val; diff_x *= roi_w;
        float diff_y = (u11 * dx + u01 * (1 - dx) - u10 * dx - u00 * (1 - dx)) * a->trans_std * diff_val; diff_y *= roi_h;
        trans_diff[tix] += diff_x; trans_diff[tiy] += diff_y;
      }
    }
  }
}

ORC_API void orc_deform_psroi_fwd(const orc_ps* a, int num_rois, const float* data, const float* rois,
                                  const float* trans, float* out, float* top_count) {
  ps_run(a, num_rois, data, rois, trans, out, top_count, 0, NULL, NULL, NULL);
}

ORC_API void orc_deform_psroi_bwd(const orc_ps* a, int num_rois, const float* out_grad, const float* data,
                                  const float* rois, const float* trans, const float* top_count, float* in_grad,
                                  float* trans_grad) {
  ps_run(a, num_rois, data, rois, trans, NULL, (float*)top_count, 1, out_grad, in_grad, trans_grad);
}
