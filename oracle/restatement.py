"""CPU restatement of IODINE's iterative-refinement inference loop.  TEST INFRASTRUCTURE.

Follows ``/root/reference/lib/modeling/iodine.py`` function by function, but WITHOUT
autograd: the five gradient tensors the reference obtains from ``(B*elbo).backward()``
(iodine.py:90) are written in closed form and the decoder
data-gradient is an explicit transposed-conv chain.  This is the executable specification
the CUDA kernels are checked against; it is validated against the unmodified reference by
``tests/test_oracle.py`` (live, when /root/reference exists) and against the committed
golden vectors (always).

Pinned: against outputs of the reference itself (see oracle/__init__.py).  The reference
holds no tests/golden vectors of its own for this path.

Weights are a plain dict keyed by the reference's state_dict names.
All math runs in the dtype of the weights (fp32 to mirror the reference, fp64 as "truth").
"""
import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- helpers
def coords_planes(H, W, dtype):
    """SpatialBroadcast coordinates (iodine.py:526-530): channel 0 = x (varies along W),
    channel 1 = y (varies along H), both linspace(-1, 1)."""
    xx = torch.linspace(-1, 1, W, dtype=dtype)
    yy = torch.linspace(-1, 1, H, dtype=dtype)
    yy, xx = torch.meshgrid((yy, xx), indexing='ij')
    return torch.stack((xx, yy), dim=0)


def _n_layers(sd, prefix):
    n = 0
    while '%s.%d.weight' % (prefix, n) in sd:
        n += 1
    return n


def layernorm3(x):
    """iodine.py:382-384,394: over the last dim, UNBIASED std, eps added to std."""
    m = x.mean(dim=2, keepdim=True)
    s = x.std(dim=2, keepdim=True)
    return (x - m) / (s + 1e-5)


def layernorm5(x):
    """iodine.py:385-394: over (C,H,W) per (b,k), BIASED std, eps added to std."""
    m = x.mean(dim=(2, 3, 4), keepdim=True)
    s = torch.sqrt(((x - m) ** 2).mean(dim=(2, 3, 4), keepdim=True))
    return (x - m) / (s + 1e-5)


# --------------------------------------------------------------------------- decoder
def decoder_forward(sd, z, img_size):
    """Decoder.forward (iodine.py:425-444) + SpatialBroadcast (512-540) + MultiLayerConv
    (586-594).  z: [B,K,L].  Returns (mean[B,K,3,H,W], logits[B,K,1,H,W], acts, out4)."""
    B, K, L = z.shape
    H = W = img_size
    zf = z.reshape(B * K, L)
    xb = zf[:, :, None, None].expand(B * K, L, H, W)
    co = coords_planes(H, W, z.dtype)[None].expand(B * K, 2, H, W)
    h = torch.cat((xb, co), dim=1)
    acts = []
    n = _n_layers(sd, 'decoder.mlc.layers')
    for i in range(n):
        w = sd['decoder.mlc.layers.%d.weight' % i]
        b = sd['decoder.mlc.layers.%d.bias' % i]
        h = F.elu(F.conv2d(h, w, b, padding=w.shape[-1] // 2))
        acts.append(h)
    wf, bf = sd['decoder.conv.weight'], sd['decoder.conv.bias']
    out4 = F.conv2d(h, wf, bf, padding=wf.shape[-1] // 2)
    mean = torch.sigmoid(out4[:, :3]).reshape(B, K, 3, H, W)
    logits = out4[:, 3:4].reshape(B, K, 1, H, W)
    return mean, logits, acts, out4


def decoder_dgrad(sd, acts, seed4, L):
    """Data-gradient of the decoder w.r.t. z (what autograd computes at iodine.py:90 for
    the decoder, minus the weight gradients nobody reads).  seed4: [BK,4,H,W] = dJ/d(out4).
    ELU'(pre) is recovered from the saved post-activation: a>0 ? 1 : a+1."""
    wf = sd['decoder.conv.weight']
    g = F.conv_transpose2d(seed4, wf, padding=wf.shape[-1] // 2)
    n = len(acts)
    for i in range(n - 1, -1, -1):
        a = acts[i]
        g = g * torch.where(a > 0, torch.ones_like(a), a + 1.0)
        w = sd['decoder.mlc.layers.%d.weight' % i]
        g = F.conv_transpose2d(g, w, padding=w.shape[-1] // 2)
    # g: [BK, L+2, H, W]; broadcast backward = sum over pixels of the first L channels
    return g[:, :L].sum(dim=(2, 3))


# --------------------------------------------------------------------------- mixture
def mixture(x, mean, logits, sigma):
    """IODINE.elbo 185-220 and the closed-form gradients of J = B*elbo (row A4).

    Returns dict with mask, K_ll[B,K,3,H,W], s[B,3,H,W], ll_sum (sum over b,c,p of s),
    mean_grad = dJ/dmean, mask_grad = dJ/dmask, seed4 = dJ/d(out4) [B*K,4,H,W]."""
    B, K = mean.shape[:2]
    mask = F.softmax(logits, dim=1)                                     # 185
    K_ll = (-(x[:, None] - mean) ** 2 / (2 * sigma ** 2)
            - math.log(sigma) - 0.5 * math.log(2 * math.pi))            # 661-666
    a = torch.log(mask + 1e-12) + K_ll                                  # 214
    s = torch.logsumexp(a, dim=1)                                       # 213-216
    r = torch.exp(a - s[:, None])
    mean_grad = r * (x[:, None] - mean) / (sigma ** 2)
    mask_grad = (r / (mask + 1e-12)).sum(dim=2, keepdim=True)
    dlogit = mask * (mask_grad - (mask * mask_grad).sum(dim=1, keepdim=True))
    dpre = mean_grad * mean * (1 - mean)
    seed4 = torch.cat((dpre, dlogit), dim=2).reshape(B * K, 4, *mean.shape[-2:])
    return dict(mask=mask, K_ll=K_ll, s=s, ll_sum=s.sum(), mean_grad=mean_grad,
                mask_grad=mask_grad, seed4=seed4)


def kl_elementwise(mu, logvar):
    """Gaussian.kl_divergence (iodine.py:653-659)."""
    return 0.5 * (torch.exp(logvar) + mu ** 2 - 1 - logvar)


# --------------------------------------------------------------------------- aux input
def input_encoding(x, mean, logits, mx, mu, logvar, mu_grad, lv_grad, layernorm=True):
    """IODINE.get_input_encoding (iodine.py:243-343), all twelve encodings, code order."""
    B, K = mean.shape[:2]
    H, W = mean.shape[-2:]
    ln3 = layernorm3 if layernorm else (lambda t: t)
    ln5 = layernorm5 if layernorm else (lambda t: t)
    latent = torch.cat((mu, logvar, ln3(mu_grad), ln3(lv_grad)), dim=-1)   # 253-275
    mask, K_ll, s = mx['mask'], mx['K_ll'], mx['s']
    Kl = torch.exp(K_ll.sum(dim=2, keepdim=True))                       # 289-290
    mask_post = Kl / Kl.sum(dim=1, keepdim=True)                        # 292 (un-stabilised)
    lik = torch.exp(s.sum(dim=1, keepdim=True))[:, None].expand(B, K, 1, H, W)  # 309-312
    tot = (mask * Kl).sum(dim=1, keepdim=True)                          # 324
    loo = (tot - mask * Kl) / (1 - mask + 1e-5)                         # 326-328
    co = coords_planes(H, W, x.dtype)[None, None].expand(B, K, 2, H, W)  # 334-339
    enc = torch.cat((
        x[:, None].expand(B, K, 3, H, W), mean, mask, logits, mask_post,
        ln5(mx['mean_grad']), ln5(mx['mask_grad']), ln5(lik), ln5(loo), co), dim=2)
    return enc, latent


# --------------------------------------------------------------------------- refinement net
def refine_forward(sd, enc, latent, hidden, stride=2):
    """RefinementNetwork.forward (iodine.py:466-503).  hidden = (h, c) by TRUE role or None.
    The reference unpacks LSTMCell's (h', c') as (c, h) (488) and feeds the variable it
    calls h -- i.e. the CELL state c' -- to both heads (491-492)."""
    B, K = enc.shape[:2]
    h = enc.reshape(B * K, *enc.shape[2:])
    n = _n_layers(sd, 'refine.mlc.layers')
    for i in range(n):
        w = sd['refine.mlc.layers.%d.weight' % i]
        b = sd['refine.mlc.layers.%d.bias' % i]
        h = F.elu(F.conv2d(h, w, b, stride=stride, padding=w.shape[-1] // 2))
    h = h.mean(dim=(2, 3))                                              # 481
    nm = _n_layers(sd, 'refine.mlp.layers')
    for i in range(nm):
        h = F.elu(F.linear(h, sd['refine.mlp.layers.%d.weight' % i],
                           sd['refine.mlp.layers.%d.bias' % i]))        # 565
    h = F.elu(h)                                                        # 485 (second ELU)
    xin = torch.cat((h, latent.reshape(B * K, -1)), dim=1)              # 487
    M = sd['refine.lstm.weight_hh'].shape[1]
    if hidden is None:
        h0 = torch.zeros(B * K, M, dtype=xin.dtype)
        c0 = torch.zeros(B * K, M, dtype=xin.dtype)
    else:
        h0, c0 = hidden
    gates = (F.linear(xin, sd['refine.lstm.weight_ih'], sd['refine.lstm.bias_ih'])
             + F.linear(h0, sd['refine.lstm.weight_hh'], sd['refine.lstm.bias_hh']))
    i_g, f_g, g_g, o_g = gates.chunk(4, dim=1)                          # torch gate order i,f,g,o
    c1 = torch.sigmoid(f_g) * c0 + torch.sigmoid(i_g) * torch.tanh(g_g)
    h1 = torch.sigmoid(o_g) * torch.tanh(c1)
    dmu = F.linear(c1, sd['refine.mean_update.weight'], sd['refine.mean_update.bias'])
    dlv = F.linear(c1, sd['refine.logvar_update.weight'], sd['refine.logvar_update.bias'])
    L = dmu.shape[1]
    return dmu.reshape(B, K, L), dlv.reshape(B, K, L), (h1, c1)


# --------------------------------------------------------------------------- the loop
def refine_step(sd, arch, x, eps_t, mu, logvar, hidden, want_aux=False):
    """One iteration of the loop body at iodine.py:83-100.  Returns new state + record."""
    B = x.shape[0]
    L = arch.DIM_LATENT
    z = mu + torch.exp(0.5 * logvar) * eps_t                            # 626-634
    mean, logits, acts, _ = decoder_forward(sd, z, arch.IMG_SIZE)
    mx = mixture(x, mean, logits, arch.SIGMA)
    kl_el = kl_elementwise(mu, logvar)
    kl = kl_el.sum() / B                                                # 193
    ll = mx['ll_sum'] / B                                               # 220
    dz = decoder_dgrad(sd, acts, mx['seed4'], L).reshape(B, arch.SLOTS, L)
    mu_grad = dz - mu
    lv_grad = dz * 0.5 * torch.exp(0.5 * logvar) * eps_t - 0.5 * (torch.exp(logvar) - 1)
    enc, latent = input_encoding(x, mean, logits, mx, mu, logvar, mu_grad, lv_grad,
                                 arch.LAYERNORM)
    dmu, dlv, hidden = refine_forward(sd, enc, latent, hidden, arch.REF.STRIDE)
    rec = dict(elbo=ll - kl, kl=kl, ll=ll, z=z, mean=mean, mask_logits=logits,
               mask=mx['mask'], mean_grad=mx['mean_grad'], mask_grad=mx['mask_grad'],
               post_mean=mu, post_logvar=logvar, post_mean_grad=mu_grad,
               post_logvar_grad=lv_grad, latent=latent, mean_delta=dmu, logvar_delta=dlv,
               lstm_h=hidden[0], lstm_c=hidden[1])
    if want_aux:
        rec['aux'] = enc
    return mu + dmu, logvar + dlv, hidden, rec                          # 642-643


def decode(sd, arch, z):
    """IODINE.decode (iodine.py:59-71)."""
    mean, logits, _, _ = decoder_forward(sd, z, arch.IMG_SIZE)
    mask = F.softmax(logits, dim=1)
    pred = (mask * mean).sum(dim=1)
    return pred, mask, mean


def encode_trace(sd, arch, x, eps, want_aux=False):
    """IODINE.encode (iodine.py:73-105) followed by decode; same record layout as
    ``oracle.ref_loader.run_reference_trace``."""
    B = x.shape[0]
    K, L = arch.SLOTS, arch.DIM_LATENT
    mu = sd['posterior.init_mean'][None, None].expand(B, K, L).clone()       # 615
    logvar = sd['posterior.init_logvar'][None, None].expand(B, K, L).clone() # 616
    hidden = None
    tr = {'steps': []}
    for t in range(arch.ITERS):
        mu, logvar, hidden, rec = refine_step(sd, arch, x, eps[t], mu, logvar, hidden, want_aux)
        tr['steps'].append(rec)
    tr['post_mean'], tr['post_logvar'] = mu, logvar
    z = mu + torch.exp(0.5 * logvar) * eps[arch.ITERS]                  # 103
    tr['z'] = z
    tr['pred'], tr['mask'], tr['mean'] = decode(sd, arch, z)
    return tr


def reconstruct(sd, arch, x, eps):
    tr = encode_trace(sd, arch, x, eps)
    return tr['pred'], tr['mask'], tr['mean'], tr


def state_dict_to(sd, dtype):
    return {k: v.detach().to(dtype).contiguous() for k, v in sd.items()}
