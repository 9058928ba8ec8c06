"""CPU restatement of IODINE's TRAINING gradients.  TEST INFRASTRUCTURE.

``IODINE.forward`` (``lib/modeling/iodine.py:115-158``) + ``loss.backward()`` (``lib/engine/train.py:60-65``), with
every parameter gradient written out explicitly -- no autograd -- so that it can serve as the executable
specification of the weight-gradient kernels.  It reuses the inference restatement (``oracle/restatement.py``) for
the forward quantities and is validated against the UNMODIFIED reference's autograd in ``tests/test_train_oracle.py``.

The structure of the reference's graph, which decides what has to be computed (and stored) at all:

* loss ``= -sum_{i=0..T} w_i * elbo_i``, ``w_i = (i+1)/(T+1)`` (iodine.py:149-153), ``elbo_i = J_i / B`` with
  ``J_i`` the batch SUM the inference loop already differentiates; ``c_i = -w_i / B`` turns every gradient of ``J_i``
  the loop computes into a gradient of the loss.
* ``Gaussian.update`` detaches the previous posterior (iodine.py:642-643) and ``get_input_encoding`` detaches both of
  its outputs (iodine.py:343).  Hence
    - the DECODER weights only receive the direct terms ``c_i * dJ_i/dW`` -- the seeds and the data-gradient chain of
      step i are exactly the inference ones, so the weight gradients can be accumulated inside each step and nothing of
      the decoder has to be kept across steps;
    - ``posterior.init_mean / init_logvar`` only see step 0: ``c_0 * sum_{b,k} dJ_0/d(posterior_0)``;
    - the REFINER is reached through ``delta_i`` (``posterior_{i+1} = const + delta_i``): its incoming gradient is
      ``c_{i+1} * dJ_{i+1}/d(posterior_{i+1})`` -- the posterior gradients the loop computes anyway (the final ELBO
      needs them too) -- plus the LSTM state chain ``(h_i, c_i) -> step i+1`` (not detached in ``forward``), i.e. a
      backward pass over the T refiner calls after the loop, with the refiner's inputs as constants (no data-gradient
      into the 17-channel stack).
* ``(B * elbo).backward(retain_graph=True)`` inside the loop (iodine.py:137) also accumulates parameter gradients,
  which ``optimizer.zero_grad()`` (train.py:62) discards before ``loss.backward()``: they are not part of the result.
"""
import torch
import torch.nn.functional as F
from torch.nn.grad import conv2d_weight

from . import restatement as S


def _elu_grad_from_act(a):
    """ELU'(pre-activation) recovered from the post-activation: 1 for a > 0, a + 1 otherwise."""
    return torch.where(a > 0, torch.ones_like(a), a + 1.0)


def _acc(g, key, val):
    g[key] = g[key] + val if key in g else val


# --------------------------------------------------------------------------- decoder
def decoder_param_grads(sd, z, acts, seed4, coef, grads):
    """accumulate ``coef * dJ/d(decoder parameters)`` for one step; returns dJ/dz [BK, L] (as decoder_dgrad)."""
    BK, L = z.shape[0] * z.shape[1], z.shape[2]
    H, W = acts[0].shape[-2:]
    zf = z.reshape(BK, L)
    h0 = torch.cat((zf[:, :, None, None].expand(BK, L, H, W), S.coords_planes(H, W, z.dtype)[None].expand(BK, 2, H, W)), dim=1)
    wf = sd['decoder.conv.weight']
    pad = wf.shape[-1] // 2
    _acc(grads, 'decoder.conv.weight', coef * conv2d_weight(acts[-1], wf.shape, seed4, padding=pad))
    _acc(grads, 'decoder.conv.bias', coef * seed4.sum(dim=(0, 2, 3)))
    g = F.conv_transpose2d(seed4, wf, padding=pad)
    for i in range(len(acts) - 1, -1, -1):
        g = g * _elu_grad_from_act(acts[i])                      # gradient w.r.t. the pre-activation of layer i
        w = sd['decoder.mlc.layers.%d.weight' % i]
        pad = w.shape[-1] // 2
        inp = acts[i - 1] if i > 0 else h0
        _acc(grads, 'decoder.mlc.layers.%d.weight' % i, coef * conv2d_weight(inp, w.shape, g, padding=pad))
        _acc(grads, 'decoder.mlc.layers.%d.bias' % i, coef * g.sum(dim=(0, 2, 3)))
        g = F.conv_transpose2d(g, w, padding=pad)
    return g[:, :L].sum(dim=(2, 3))


# --------------------------------------------------------------------------- refiner, forward with a tape
def refine_forward_tape(sd, enc, latent, hidden, stride=2):
    """``restatement.refine_forward`` that also returns what the backward pass needs."""
    B, K = enc.shape[:2]
    t = {'inputs': [], 'acts': []}
    h = enc.reshape(B * K, *enc.shape[2:])
    n = S._n_layers(sd, 'refine.mlc.layers')
    for i in range(n):
        w, b = sd['refine.mlc.layers.%d.weight' % i], sd['refine.mlc.layers.%d.bias' % i]
        t['inputs'].append(h)
        h = F.elu(F.conv2d(h, w, b, stride=stride, padding=w.shape[-1] // 2))
        t['acts'].append(h)
    t['pool_hw'] = h.shape[-2:]
    h = h.mean(dim=(2, 3))
    t['mlp_in'], t['mlp_act'] = [], []
    for i in range(S._n_layers(sd, 'refine.mlp.layers')):
        t['mlp_in'].append(h)
        h = F.elu(F.linear(h, sd['refine.mlp.layers.%d.weight' % i], sd['refine.mlp.layers.%d.bias' % i]))
        t['mlp_act'].append(h)
    t['pre_elu2'] = h
    h = F.elu(h)
    xin = torch.cat((h, latent.reshape(B * K, -1)), dim=1)
    M = sd['refine.lstm.weight_hh'].shape[1]
    h0, c0 = hidden if hidden is not None else (torch.zeros(B * K, M, dtype=xin.dtype), torch.zeros(B * K, M, dtype=xin.dtype))
    gates = (F.linear(xin, sd['refine.lstm.weight_ih'], sd['refine.lstm.bias_ih'])
             + F.linear(h0, sd['refine.lstm.weight_hh'], sd['refine.lstm.bias_hh']))
    i_g, f_g, g_g, o_g = gates.chunk(4, dim=1)
    c1 = torch.sigmoid(f_g) * c0 + torch.sigmoid(i_g) * torch.tanh(g_g)
    h1 = torch.sigmoid(o_g) * torch.tanh(c1)
    dmu = F.linear(c1, sd['refine.mean_update.weight'], sd['refine.mean_update.bias'])
    dlv = F.linear(c1, sd['refine.logvar_update.weight'], sd['refine.logvar_update.bias'])
    t.update(xin=xin, h0=h0, c0=c0, gates=gates, c1=c1, n_feat=h.shape[1])
    L = dmu.shape[1]
    return dmu.reshape(B, K, L), dlv.reshape(B, K, L), (h1, c1), t


def refine_backward(sd, t, g_dmu, g_dlv, dh_next, dc_next, grads, stride=2):
    """backward of one refiner call.  g_dmu / g_dlv: [BK, L] gradients of the loss w.r.t. this call's deltas;
    dh_next / dc_next: [BK, M] gradients arriving through the LSTM state from the following call (or zeros).
    Accumulates the parameter gradients into ``grads`` and returns (dh_prev, dc_prev)."""
    c1, c0, h0, xin = t['c1'], t['c0'], t['h0'], t['xin']
    i_g, f_g, g_g, o_g = t['gates'].chunk(4, dim=1)
    si, sf, so, tg, tc = torch.sigmoid(i_g), torch.sigmoid(f_g), torch.sigmoid(o_g), torch.tanh(g_g), torch.tanh(c1)
    # heads read the CELL state (iodine.py:488-492)
    Wm, Wl = sd['refine.mean_update.weight'], sd['refine.logvar_update.weight']
    _acc(grads, 'refine.mean_update.weight', g_dmu.t() @ c1)
    _acc(grads, 'refine.mean_update.bias', g_dmu.sum(0))
    _acc(grads, 'refine.logvar_update.weight', g_dlv.t() @ c1)
    _acc(grads, 'refine.logvar_update.bias', g_dlv.sum(0))
    dc = dc_next + g_dmu @ Wm + g_dlv @ Wl
    dh = dh_next
    # LSTM cell (gate order i, f, g, o)
    do = dh * tc * so * (1 - so)
    dc = dc + dh * so * (1 - tc * tc)
    di = dc * tg * si * (1 - si)
    df = dc * c0 * sf * (1 - sf)
    dg = dc * si * (1 - tg * tg)
    dgates = torch.cat((di, df, dg, do), dim=1)
    _acc(grads, 'refine.lstm.weight_ih', dgates.t() @ xin)
    _acc(grads, 'refine.lstm.weight_hh', dgates.t() @ h0)
    _acc(grads, 'refine.lstm.bias_ih', dgates.sum(0))
    _acc(grads, 'refine.lstm.bias_hh', dgates.sum(0))
    dh_prev = dgates @ sd['refine.lstm.weight_hh']
    dc_prev = dc * sf
    d = (dgates @ sd['refine.lstm.weight_ih'])[:, :t['n_feat']]        # the latent half of xin is detached (iodine.py:343)
    # second ELU (iodine.py:485), then the MLP (each layer ELU(Linear), iodine.py:565)
    u = t['pre_elu2']
    d = d * torch.where(u > 0, torch.ones_like(u), torch.exp(u))
    for i in range(len(t['mlp_act']) - 1, -1, -1):
        d = d * _elu_grad_from_act(t['mlp_act'][i])
        _acc(grads, 'refine.mlp.layers.%d.weight' % i, d.t() @ t['mlp_in'][i])
        _acc(grads, 'refine.mlp.layers.%d.bias' % i, d.sum(0))
        d = d @ sd['refine.mlp.layers.%d.weight' % i]
    # adaptive average pool (iodine.py:481)
    ph, pw = t['pool_hw']
    g = (d / (ph * pw))[:, :, None, None].expand(-1, -1, ph, pw)
    for i in range(len(t['acts']) - 1, -1, -1):
        g = g * _elu_grad_from_act(t['acts'][i])
        w = sd['refine.mlc.layers.%d.weight' % i]
        k, pad = w.shape[-1], w.shape[-1] // 2
        inp = t['inputs'][i]
        _acc(grads, 'refine.mlc.layers.%d.weight' % i, conv2d_weight(inp, w.shape, g, stride=stride, padding=pad))
        _acc(grads, 'refine.mlc.layers.%d.bias' % i, g.sum(dim=(0, 2, 3)))
        if i > 0:                                                     # the 17-channel input itself is a constant
            op = tuple(inp.shape[-2 + a] - ((g.shape[-2 + a] - 1) * stride - 2 * pad + k) for a in (0, 1))
            g = F.conv_transpose2d(g, w, stride=stride, padding=pad, output_padding=op)
    return dh_prev, dc_prev


# --------------------------------------------------------------------------- the training step
def loss_and_grads(sd, arch, x, eps, global_batch=None):
    """``-weighted ELBO`` of ``IODINE.forward`` and its gradient w.r.t. every entry of the state_dict.
    x: [B,3,H,W]; eps: [T+1,B,K,L] (the T+1 ``torch.randn_like`` draws).  Returns (loss, grads, elbos[T+1]).

    ``global_batch``: slot-shard data parallelism (SURVEY.md 8e "Training"): x holds only this rank's images of a
    batch of ``global_batch``; every batch mean divides by the GLOBAL size, so loss, ELBOs and gradients are this
    rank's additive share and ONE sum all-reduce of the gradients (4.4 MB) gives the full-batch result -- nothing
    else crosses ranks, the loop and both backward passes are per image."""
    B, K, L, T = x.shape[0], arch.SLOTS, arch.DIM_LATENT, arch.ITERS
    Bg = B if global_batch is None else int(global_batch)
    grads, tapes, post_grads, elbos = {}, [], [], []
    mu = sd['posterior.init_mean'][None, None].expand(B, K, L).clone()
    lv = sd['posterior.init_logvar'][None, None].expand(B, K, L).clone()
    hidden = None
    for i in range(T + 1):
        coef = -((i + 1) / (T + 1)) / Bg
        z = mu + torch.exp(0.5 * lv) * eps[i]
        mean, logits, acts, _ = S.decoder_forward(sd, z, arch.IMG_SIZE)
        mx = S.mixture(x, mean, logits, arch.SIGMA)
        elbos.append((mx['ll_sum'] - S.kl_elementwise(mu, lv).sum()) / Bg)
        dz = decoder_param_grads(sd, z, acts, mx['seed4'], coef, grads).reshape(B, K, L)
        mu_grad = dz - mu                                              # dJ_i / d(posterior_i), as in the loop
        lv_grad = dz * 0.5 * torch.exp(0.5 * lv) * eps[i] - 0.5 * (torch.exp(lv) - 1)
        post_grads.append((coef * mu_grad, coef * lv_grad))
        if i == 0:                                                     # init_unit: repeat over (B, K) (iodine.py:615-616)
            grads['posterior.init_mean'] = (coef * mu_grad).sum(dim=(0, 1))
            grads['posterior.init_logvar'] = (coef * lv_grad).sum(dim=(0, 1))
        if i == T:
            break
        enc, latent = S.input_encoding(x, mean, logits, mx, mu, lv, mu_grad, lv_grad, arch.LAYERNORM)
        dmu, dlv, hidden, tape = refine_forward_tape(sd, enc, latent, hidden, arch.REF.STRIDE)
        tapes.append(tape)
        mu, lv = mu + dmu, lv + dlv                                    # detach(prev) + delta (iodine.py:642-643)
    M = sd['refine.lstm.weight_hh'].shape[1]
    dh = torch.zeros(B * K, M, dtype=x.dtype)
    dc = torch.zeros(B * K, M, dtype=x.dtype)
    for i in range(T - 1, -1, -1):
        g_mu, g_lv = post_grads[i + 1]
        dh, dc = refine_backward(sd, tapes[i], g_mu.reshape(B * K, L), g_lv.reshape(B * K, L), dh, dc, grads,
                                 arch.REF.STRIDE)
    loss = -sum((i + 1) / (T + 1) * e for i, e in enumerate(elbos))
    return loss, grads, torch.stack(elbos)
