"""Checkpoint loading: the reference's pretrained-weight entry points without its download step.

reference: models/utils.py:11-197 (``load_weights_from_npz``, Google's Flax ViT checkpoints),
models/vit_fquant.py:820-932 (which checkpoint each factory takes), utils/build_model.py:64-93 (``build_model``).

This image has no network, so nothing here downloads.  A checkpoint is taken from a local path, or from the torch hub
cache (``torch.hub.get_dir()/checkpoints/<file name of the reference's URL>``) where the reference's own download would
have left it; if it is not there the error names the file that is missing.
"""
import os

import numpy as np
import torch
import torch.nn.functional as F

# file names of the URLs the reference's factories download (models/vit_fquant.py:822-932)
CHECKPOINTS = {
    'deit_tiny': 'deit_tiny_patch16_224-a1311bcf.pth',
    'deit_small': 'deit_small_patch16_224-cd65a155.pth',
    'deit_base': 'deit_base_patch16_224-b5f2ef4d.pth',
    'vit_base': 'B_16-i21k-300ep-lr_0.001-aug_medium1-wd_0.1-do_0.0-sd_0.0--imagenet2012-steps_20k-lr_0.01-res_224.npz',
    'vit_large': 'L_16-i21k-300ep-lr_0.001-aug_medium1-wd_0.1-do_0.1-sd_0.1--imagenet2012-steps_20k-lr_0.01-res_224.npz',
}


def cached_checkpoint(file_name):
    """Path of `file_name` in the torch hub checkpoint cache (where torch.hub.load_state_dict_from_url and the
    reference's _download_cached_file put their downloads); RuntimeError if it is not there."""
    path = os.path.join(torch.hub.get_dir(), 'checkpoints', file_name)
    if not os.path.exists(path):
        raise RuntimeError('pretrained checkpoint %s is not in the torch hub cache (%s) and this build never downloads: '
                           'copy the file there, or pass its path as pretrained=<path>' % (file_name, os.path.dirname(path)))
    return path


def _to_torch(w, transpose=True):
    """Flax kernel -> torch weight: HWIO -> OIHW, [in, out] -> [out, in] (reference: models/utils.py:20-30)."""
    if w.ndim == 4 and w.shape[0] == w.shape[1] == w.shape[2] == 1:
        w = w.flatten()
    if transpose:
        if w.ndim == 4:
            w = w.transpose([3, 2, 0, 1])
        elif w.ndim == 3:
            w = w.transpose([2, 0, 1])
        elif w.ndim == 2:
            w = w.transpose([1, 0])
    return torch.from_numpy(np.ascontiguousarray(w))


def resize_pos_embed(posemb, ntok_new, num_tokens=1, gs_new=()):
    """Bicubic resize of the grid part of a position embedding (reference: models/utils.py:86-110)."""
    if num_tokens:
        tok, grid = posemb[:, :num_tokens], posemb[0, num_tokens:]
        ntok_new -= num_tokens
    else:
        tok, grid = posemb[:, :0], posemb[0]
    gs_old = int(round(len(grid) ** 0.5))
    if not len(gs_new):
        gs_new = [int(round(ntok_new ** 0.5))] * 2
    grid = grid.reshape(1, gs_old, gs_old, -1).permute(0, 3, 1, 2)
    grid = F.interpolate(grid, size=tuple(gs_new), mode='bicubic', align_corners=False)
    grid = grid.permute(0, 2, 3, 1).reshape(1, gs_new[0] * gs_new[1], -1)
    return torch.cat([tok, grid], dim=1)


def _adapt_input_conv(in_chans, w):
    """RGB stem kernel -> `in_chans` input channels (reference: models/utils.py:58-84)."""
    if in_chans == 3:
        return w
    dtype, w = w.dtype, w.float()
    if in_chans == 1:
        w = w.sum(dim=1, keepdim=True)
    else:
        if w.shape[1] != 3:
            raise NotImplementedError('stem kernel with %d input channels cannot be adapted' % w.shape[1])
        repeat = -(-in_chans // 3)
        w = w.repeat(1, repeat, 1, 1)[:, :in_chans] * (3.0 / in_chans)
    return w.to(dtype)


@torch.no_grad()
def load_weights_from_npz(model, path, check_hash=False, progress=False, prefix=''):
    """Copy a Flax ViT checkpoint (.npz, the layout of google-research/vision_transformer) into `model`.

    Same call as the reference's (models/utils.py:11-197) except that `path` is a local file, or a URL whose file name is
    looked up in the torch hub cache; `check_hash` and `progress` are accepted for signature compatibility."""
    if '://' in path:
        path = cached_checkpoint(os.path.basename(path))
    w = np.load(path)
    if not prefix and 'opt/target/embedding/kernel' in w:
        prefix = 'opt/target/'
    if hasattr(model.patch_embed, 'backbone'):
        raise NotImplementedError('hybrid (ResNet stem) checkpoints are outside this package: no model here has one')

    def get(name, transpose=True):
        return _to_torch(w[prefix + name], transpose)

    conv = model.patch_embed.proj
    conv.weight.copy_(_adapt_input_conv(conv.weight.shape[1], get('embedding/kernel')))
    conv.bias.copy_(get('embedding/bias'))
    model.cls_token.copy_(get('cls', False))
    pos = get('Transformer/posembed_input/pos_embedding', False)
    if pos.shape != model.pos_embed.shape:
        pos = resize_pos_embed(pos, model.pos_embed.shape[1], getattr(model, 'num_tokens', 1), model.patch_embed.grid_size)
    model.pos_embed.copy_(pos)
    model.norm.weight.copy_(get('Transformer/encoder_norm/scale'))
    model.norm.bias.copy_(get('Transformer/encoder_norm/bias'))
    head = getattr(model, 'head', None)
    if isinstance(head, torch.nn.Linear) and head.bias.shape[0] == w[prefix + 'head/bias'].shape[-1]:
        head.weight.copy_(get('head/kernel'))
        head.bias.copy_(get('head/bias'))
    fc = getattr(getattr(model, 'pre_logits', None), 'fc', None)
    if isinstance(fc, torch.nn.Linear) and prefix + 'pre_logits/bias' in w:
        fc.weight.copy_(get('pre_logits/kernel'))
        fc.bias.copy_(get('pre_logits/bias'))
    for i, block in enumerate(model.blocks.children()):
        bp = 'Transformer/encoderblock_%d/' % i
        mha = bp + 'MultiHeadDotProductAttention_1/'
        block.norm1.weight.copy_(get(bp + 'LayerNorm_0/scale'))
        block.norm1.bias.copy_(get(bp + 'LayerNorm_0/bias'))
        # Flax keeps q / k / v as [in, heads, head_dim]: flatten to [in, out], transpose, stack
        block.attn.qkv.weight.copy_(torch.cat([get(mha + n + '/kernel', False).flatten(1).T for n in ('query', 'key', 'value')]))
        block.attn.qkv.bias.copy_(torch.cat([get(mha + n + '/bias', False).reshape(-1) for n in ('query', 'key', 'value')]))
        block.attn.proj.weight.copy_(get(mha + 'out/kernel').flatten(1))
        block.attn.proj.bias.copy_(get(mha + 'out/bias'))
        for r in range(2):
            fc_r = getattr(block.mlp, 'fc%d' % (r + 1))
            fc_r.weight.copy_(get(bp + 'MlpBlock_3/Dense_%d/kernel' % r))
            fc_r.bias.copy_(get(bp + 'MlpBlock_3/Dense_%d/bias' % r))
        block.norm2.weight.copy_(get(bp + 'LayerNorm_2/scale'))
        block.norm2.bias.copy_(get(bp + 'LayerNorm_2/bias'))
    return model


@torch.no_grad()
def export_weights_to_npz(model, path, prefix=''):
    """The inverse of load_weights_from_npz: write `model`'s float weights in the Flax ViT layout.  Lets a user move
    weights between the two ecosystems offline, and gives the loader a round-trip test that needs no download."""
    out = {}

    def put(name, t, transpose=True):
        a = t.detach().cpu().numpy()
        if transpose:
            if a.ndim == 4:
                a = a.transpose([2, 3, 1, 0])   # OIHW -> HWIO
            elif a.ndim == 2:
                a = a.transpose([1, 0])
        out[prefix + name] = np.ascontiguousarray(a)

    heads = model.blocks[0].attn.num_heads
    put('embedding/kernel', model.patch_embed.proj.weight)
    put('embedding/bias', model.patch_embed.proj.bias)
    put('cls', model.cls_token, False)
    put('Transformer/posembed_input/pos_embedding', model.pos_embed, False)
    put('Transformer/encoder_norm/scale', model.norm.weight)
    put('Transformer/encoder_norm/bias', model.norm.bias)
    if isinstance(getattr(model, 'head', None), torch.nn.Linear):
        put('head/kernel', model.head.weight)
        put('head/bias', model.head.bias)
    for i, block in enumerate(model.blocks.children()):
        bp = 'Transformer/encoderblock_%d/' % i
        mha = bp + 'MultiHeadDotProductAttention_1/'
        put(bp + 'LayerNorm_0/scale', block.norm1.weight)
        put(bp + 'LayerNorm_0/bias', block.norm1.bias)
        d = block.attn.qkv.weight.shape[1]
        for j, n in enumerate(('query', 'key', 'value')):
            wj = block.attn.qkv.weight[j * d:(j + 1) * d]             # [out, in]
            put(mha + n + '/kernel', wj.T.reshape(d, heads, d // heads), False)
            put(mha + n + '/bias', block.attn.qkv.bias[j * d:(j + 1) * d].reshape(heads, d // heads), False)
        put(mha + 'out/kernel', block.attn.proj.weight.T.reshape(heads, d // heads, d), False)
        put(mha + 'out/bias', block.attn.proj.bias)
        for r in range(2):
            fc_r = getattr(block.mlp, 'fc%d' % (r + 1))
            put(bp + 'MlpBlock_3/Dense_%d/kernel' % r, fc_r.weight)
            put(bp + 'MlpBlock_3/Dense_%d/bias' % r, fc_r.bias)
        put(bp + 'LayerNorm_2/scale', block.norm2.weight)
        put(bp + 'LayerNorm_2/bias', block.norm2.bias)
    np.savez(path, **out)
    return path


@torch.no_grad()
def load_pretrained(model, short_name, pretrained):
    """What the reference's factories do under ``pretrained=True`` (models/vit_fquant.py:820-932), from local files.

    pretrained: True -> the reference's checkpoint for `short_name`, from the torch hub cache; a path -> that file
    (``.npz`` = Flax layout, anything else = a torch checkpoint whose 'model' entry, or itself, is the state_dict)."""
    if pretrained is True:
        if short_name not in CHECKPOINTS:
            raise RuntimeError('the reference names no pretrained checkpoint for %r' % short_name)
        path = cached_checkpoint(CHECKPOINTS[short_name])
    else:
        path = os.fspath(pretrained)
        if not os.path.exists(path):
            raise RuntimeError('checkpoint %s does not exist' % path)
    if path.endswith('.npz'):
        return load_weights_from_npz(model, path)
    ckpt = torch.load(path, map_location='cpu', weights_only=True)
    state = ckpt['model'] if isinstance(ckpt, dict) and 'model' in ckpt else ckpt
    model.load_state_dict(state, strict=False)   # strict=False as in the reference: quantizer buffers are not in the file
    return model


# timm names the reference's build_model is called with -> this package's factories
_TIMM_NAMES = {
    'deit_tiny_patch16_224': 'deit_tiny', 'deit_small_patch16_224': 'deit_small', 'deit_base_patch16_224': 'deit_base',
    'vit_base_patch16_224': 'vit_base', 'vit_large_patch16_224': 'vit_large',
    'swin_tiny_patch4_window7_224': 'swin_tiny', 'swin_small_patch4_window7_224': 'swin_small',
    'swin_base_patch4_window7_224': 'swin_base',
}


def build_model(name, Pretrained=True):
    """Entry-point alias of the reference's ``utils.build_model.build_model`` (utils/build_model.py:64-93).

    The reference asks timm for the float network and swaps its attention matmuls for modules; here the quantizable
    model IS that float network until it is calibrated (``quant=False``: every Q* layer passes through), so the alias
    returns the package's own model for the timm name, in eval mode, on the GPU when there is one."""
    from . import Config, str2model
    short = _TIMM_NAMES.get(name, name)
    try:
        factory = str2model(short)
    except KeyError:
        raise ValueError('build_model: no model named %r (known: %s)' % (name, ', '.join(sorted(_TIMM_NAMES)))) from None
    net = factory(pretrained=False, cfg=Config(True, True, 'minmax'))   # the reference's default quantization recipe
    if Pretrained:
        load_pretrained(net, short, Pretrained)
    if torch.cuda.is_available():
        net = net.cuda()
    net.eval()
    return net
