// pinn_host_plan.h — host-side validation of a PinnSpec and construction of the device plan
// (layer table, shared-memory weight layout, per-point storage rows).  Plain C++: included by
// pinn_kernels.cu and by the test-only emulation harness (tests/emul/pinn_emul.cpp).
#pragma once

#include <stdio.h>
#include <string.h>
#include "pinn_device.cuh"

namespace pinn {

inline int round_up_i(int x, int m) { return (x + m - 1) / m * m; }

#define PINN_PLAN_FAIL(code, ...)                   \
    do {                                            \
        snprintf(msg, msg_len, __VA_ARGS__);        \
        return (code);                              \
    } while (0)

inline int validate_prog(const PinnInstr* prog, int n, int n_slots, int total, int n_vars, const char* name,
                         char* msg, size_t msg_len) {
    if (n < 0 || n > PINN_MAX_PROG) PINN_PLAN_FAIL(PINN_E_INVALID, "%s: %d instructions (max %d)", name, n, PINN_MAX_PROG);
    for (int i = 0; i < n; ++i) {
        const PinnInstr& in = prog[i];
        if (in.op >= PINN_OP_COUNT_) PINN_PLAN_FAIL(PINN_E_INVALID, "%s[%d]: bad opcode %d", name, i, in.op);
        if (in.dst >= n_slots) PINN_PLAN_FAIL(PINN_E_INVALID, "%s[%d]: dst slot %d >= n_slots %d", name, i, in.dst, n_slots);
        if (in.op == PINN_OP_COORD && in.a >= total) PINN_PLAN_FAIL(PINN_E_INVALID, "%s[%d]: coord %d", name, i, in.a);
        if (in.op == PINN_OP_VAR && in.a >= n_vars) PINN_PLAN_FAIL(PINN_E_INVALID, "%s[%d]: var %d", name, i, in.a);
        if (in.op >= PINN_OP_ADD) {
            if (in.a >= n_slots) PINN_PLAN_FAIL(PINN_E_INVALID, "%s[%d]: src slot", name, i);
            bool binary = in.op == PINN_OP_ADD || in.op == PINN_OP_SUB || in.op == PINN_OP_MUL ||
                          in.op == PINN_OP_DIV || in.op == PINN_OP_POW;
            if (binary && in.b >= n_slots) PINN_PLAN_FAIL(PINN_E_INVALID, "%s[%d]: src slot", name, i);
        }
    }
    return PINN_OK;
}

// Derivative order of a spec: 2 (value / first / second channels), or 3 / 4 (whole jets per direction, pinn_device_hi.cuh).
inline int spec_order(const PinnSpec* s) { return s->order >= 3 ? s->order : 2; }
inline int spec_channels(const PinnSpec* s) { return s->order >= 3 ? 1 + s->nf * s->order : 1 + s->nf + s->ns; }

// Fills `h` (device plan, minus anything that needs a device) and the forward-only row layout.
inline int build_dev_plan(const PinnSpec* s, DevPlan& h, int& fwd_rows, int& fwd_row_scr, char* msg, size_t msg_len) {
    if (s->abi_version != PINN_ABI_VERSION)
        PINN_PLAN_FAIL(PINN_E_INVALID, "spec abi_version %d != library %d", s->abi_version, PINN_ABI_VERSION);
    const int Ln = s->n_layers;
    if (Ln < 1 || Ln > PINN_MAX_LAYERS) PINN_PLAN_FAIL(PINN_E_INVALID, "n_layers %d", Ln);
    const int total = s->ndims + s->nparams;
    if (s->ndims < 1 || s->nparams < 0 || total > PINN_MAX_DIMS) PINN_PLAN_FAIL(PINN_E_INVALID, "ndims/nparams");
    if (s->widths[0] != total) PINN_PLAN_FAIL(PINN_E_INVALID, "widths[0]=%d != ndims+nparams=%d", s->widths[0], total);
    if (s->widths[Ln] != 1) PINN_PLAN_FAIL(PINN_E_UNSUPPORTED, "output width %d (must be 1)", s->widths[Ln]);
    if (s->act[Ln - 1] != PINN_ACT_NONE) PINN_PLAN_FAIL(PINN_E_UNSUPPORTED, "activation after the last layer");
    if (s->n_params <= 0 || (s->n_params & 3)) PINN_PLAN_FAIL(PINN_E_INVALID, "n_params must be a positive multiple of 4");
    if (s->nf < 0 || s->nf > PINN_MAX_DIRS || s->ns < 0 || s->ns > s->nf) PINN_PLAN_FAIL(PINN_E_INVALID, "jet set nf=%d ns=%d", s->nf, s->ns);
    if (s->n_vars < 0 || s->n_vars > PINN_MAX_VARS) PINN_PLAN_FAIL(PINN_E_INVALID, "n_vars");
    if (s->order != 0 && (s->order < 2 || s->order > 4)) PINN_PLAN_FAIL(PINN_E_INVALID, "derivative order %d (2..4)", s->order);
    const bool hi = s->order >= 3;
    if (hi) {
        // whole-jet plans (pinn_device_hi.cuh): what that code covers
        if (s->nf < 1 || s->nf > 4) PINN_PLAN_FAIL(PINN_E_UNSUPPORTED, "order %d with %d derivative directions (1..4)", s->order, s->nf);
        if (s->ns != 0) PINN_PLAN_FAIL(PINN_E_INVALID, "order %d: ns must be 0 (every direction carries its whole jet)", s->order);
        for (int d = 0; d < s->nf; ++d)
            for (int e = 0; e < d; ++e)
                if (s->dir_col[d] >= 0 && s->dir_col[e] == s->dir_col[d])
                    PINN_PLAN_FAIL(PINN_E_INVALID, "order %d: directions %d and %d coincide", s->order, e, d);
    }
    const int C = spec_channels(s);
    if (1 + C + s->n_vars > (int)(sizeof(s->eq_out) / sizeof(s->eq_out[0])))
        PINN_PLAN_FAIL(PINN_E_UNSUPPORTED, "%d jet channels + %d variables exceed the program's outputs", C, s->n_vars);
    if (s->has_ic && C * (1 + (s->ic_has_vars ? s->n_vars : 0)) > (int)(sizeof(s->ic_out) / sizeof(s->ic_out[0])))
        PINN_PLAN_FAIL(PINN_E_UNSUPPORTED, "%d jet channels x (1 + %d variables) exceed the initial condition's outputs", C, s->n_vars);
    for (int d = 0; d < s->nf; ++d) {
        if (s->dir_col[d] >= total) PINN_PLAN_FAIL(PINN_E_INVALID, "dir_col[%d]", d);
        bool any = false;
        for (int k = 0; k < total; ++k) {
            const float v = s->dir_vec[d][k];
            if (!(v == v) || v > 1e6f || v < -1e6f) PINN_PLAN_FAIL(PINN_E_INVALID, "dir_vec[%d][%d]", d, k);
            any = any || v != 0.0f;
            if (s->dir_col[d] >= 0 && v != (k == s->dir_col[d] ? 1.0f : 0.0f))
                PINN_PLAN_FAIL(PINN_E_INVALID, "dir_vec[%d] is not the unit vector of column %d", d, s->dir_col[d]);
        }
        if (!any) PINN_PLAN_FAIL(PINN_E_INVALID, "dir_vec[%d] is zero", d);
    }
    if (s->n_slots < C || s->n_slots > PINN_MAX_SLOTS) PINN_PLAN_FAIL(PINN_E_INVALID, "n_slots %d", s->n_slots);
    int rc;
    if ((rc = validate_prog(s->eq_prog, s->n_eq, s->n_slots, total, s->n_vars, "eq_prog", msg, msg_len))) return rc;
    if (s->has_ic && (rc = validate_prog(s->ic_prog, s->n_ic, s->n_slots, total, s->n_vars, "ic_prog", msg, msg_len))) return rc;
    for (int i = 0; i < 1 + C + s->n_vars; ++i)
        if (s->eq_out[i] < 0 || s->eq_out[i] >= s->n_slots) PINN_PLAN_FAIL(PINN_E_INVALID, "eq_out[%d]", i);
    if (s->has_ic)
        for (int c = 0; c < C * (1 + (s->ic_has_vars ? s->n_vars : 0)); ++c)
            if (s->ic_out[c] < 0 || s->ic_out[c] >= s->n_slots) PINN_PLAN_FAIL(PINN_E_INVALID, "ic_out[%d]", c);
    if (s->log_scale_off < 0 || s->log_scale_off >= s->n_params) PINN_PLAN_FAIL(PINN_E_INVALID, "log_scale_off");
    for (int i = 0; i < s->n_vars; ++i)
        if (s->var_off[i] < 0 || s->var_off[i] >= s->n_params) PINN_PLAN_FAIL(PINN_E_INVALID, "var_off[%d]", i);

    memset(&h, 0, sizeof(h));
    h.n_layers = Ln; h.total = total; h.ndims = s->ndims; h.nparams = s->nparams;
    h.has_bc = s->has_bc ? 1 : 0; h.has_ic = s->has_ic ? 1 : 0;
    h.nsp = s->has_ic ? s->ndims - 1 : s->ndims;
    h.nf = s->nf; h.ns = hi ? 0 : s->ns; h.n_params = s->n_params; h.log_scale_off = s->log_scale_off;
    h.n_vars = s->n_vars; h.n_eq = s->n_eq; h.n_ic = s->has_ic ? s->n_ic : 0; h.n_slots = s->n_slots;
    h.bc = s->bc_value;
    h.ic_has_vars = (s->has_ic && s->ic_has_vars) ? 1 : 0;
    h.t0 = s->dom_lo[s->ndims - 1];
    for (int d = 0; d < PINN_MAX_DIRS; ++d) {
        h.dir_col[d] = d < s->nf ? s->dir_col[d] : 0;
        for (int k = 0; k < PINN_MAX_DIMS; ++k) h.dir_vec[d][k] = (d < s->nf && k < total) ? s->dir_vec[d][k] : 0.0f;
    }
    for (int i = 0; i < PINN_MAX_VARS; ++i) h.var_off[i] = i < s->n_vars ? s->var_off[i] : 0;
    for (int i = 0; i < PINN_MAX_DIMS; ++i) {
        h.lo[i] = s->dom_lo[i]; h.hi[i] = s->dom_hi[i];
        float w = s->dom_hi[i] - s->dom_lo[i];
        h.inv_w2[i] = (i < s->ndims && w != 0.0f) ? 1.0f / (w * w) : 0.0f;
        memset(&h.cols[i], 0, sizeof(PinnColumn));
        h.cols[i].kind = PINN_COL_UNIFORM; h.cols[i].a = 0.0f; h.cols[i].b = 1.0f;
    }
    memcpy(h.eq_out, s->eq_out, sizeof(h.eq_out));
    memcpy(h.ic_out, s->ic_out, sizeof(h.ic_out));
    memcpy(h.eq, s->eq_prog, sizeof(h.eq));
    memcpy(h.ic, s->ic_prog, sizeof(h.ic));

    int sw = 0, units = 0;
    for (int l = 0; l < Ln; ++l) {
        DevLayer& L = h.layer[l];
        L.n_in = s->widths[l]; L.n_out = s->widths[l + 1]; L.act = s->act[l];
        if (L.n_in < 1 || L.n_out < 1) PINN_PLAN_FAIL(PINN_E_INVALID, "layer %d width", l);
        if (L.act < 0 || L.act > PINN_ACT_GELU) PINN_PLAN_FAIL(PINN_E_UNSUPPORTED, "layer %d: activation %d is not covered by the fused kernel", l, L.act);
        L.w_off = s->w_off[l]; L.b_off = s->b_off[l];
        if (L.w_off < 0 || L.w_off + L.n_in * L.n_out > s->n_params || L.b_off < 0 || L.b_off + L.n_out > s->n_params)
            PINN_PLAN_FAIL(PINN_E_INVALID, "layer %d offsets out of range", l);
        L.n_out_p4 = round_up_i(L.n_out, 4);
        L.n_in_p8 = round_up_i(L.n_in + 1, 8);          // + the (zero) bias column of the reverse sweep
        L.wt_s = sw; sw += L.n_in * L.n_out_p4;
        L.w_s = sw;  sw += L.n_out_p4 * L.n_in_p8;
        L.b_s = sw;  sw += L.n_out_p4;
        L.unit_base = units; units += L.n_out;
        L.post_base = -1; L.skip_src = -1; L.adj_from = -1;
    }
    for (int l = 0; l < Ln; ++l) {
        const int src = s->skip_src[l];
        if (src < 0) continue;
        DevLayer& L = h.layer[l];
        if (l == Ln - 1) PINN_PLAN_FAIL(PINN_E_UNSUPPORTED, "skip connection into the output layer");
        if (src >= l) PINN_PLAN_FAIL(PINN_E_INVALID, "skip_src[%d] = %d must be an earlier layer", l, src);
        if (h.layer[src].n_out != L.n_out) PINN_PLAN_FAIL(PINN_E_INVALID, "skip %d -> %d: widths %d and %d differ", src, l, h.layer[src].n_out, L.n_out);
        if (h.layer[src].adj_from >= 0) PINN_PLAN_FAIL(PINN_E_UNSUPPORTED, "layer %d is the source of two skip connections", src);
        L.skip_src = src; L.post_base = units; units += L.n_out;
        h.layer[src].adj_from = l;
    }
    h.general = h.ic_has_vars;
    for (int l = 0; l < Ln; ++l) if (h.layer[l].skip_src >= 0) h.general = 1;
    for (int l = 0; l < Ln; ++l) if (h.layer[l].act >= PINN_ACT_SIN) h.general = 1;     // z-stored activations
    for (int d = 0; d < s->nf; ++d) if (s->dir_col[d] < 0) h.general = 1;
    h.weights_floats = round_up_i(sw, 4);
    h.n_units = units;
    h.row_units = total;
    h.row_scr = total + units * C;
    h.rows_total = h.row_scr + s->n_slots;
    fwd_row_scr = total + units;
    fwd_rows = fwd_row_scr + s->n_slots;
    return PINN_OK;
}

// Re-layout of the flat parameters into the weight area (what stage_plan_and_weights does on the GPU).
inline void host_stage_weights(const DevPlan& P, const float* params, float* sw) {
    for (int i = 0; i < P.weights_floats; ++i) sw[i] = 0.0f;
    for (int l = 0; l < P.n_layers; ++l) {
        const DevLayer& L = P.layer[l];
        for (int j = 0; j < L.n_out; ++j) {
            for (int k = 0; k < L.n_in; ++k) {
                float w = params[L.w_off + j * L.n_in + k];
                sw[L.wt_s + k * L.n_out_p4 + j] = w;
                sw[L.w_s + j * L.n_in_p8 + k] = w;
            }
            sw[L.b_s + j] = params[L.b_off + j];
        }
    }
}

}  // namespace pinn
