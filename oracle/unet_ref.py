"""fp32 PyTorch restatement of the reference U-Net forward.  TEST INFRASTRUCTURE.

Functional (state_dict in, tensor out) restatement of
  UNet.forward / encoder / decoder   (src/iterseg/unet.py:284-364)
  ConvModule.forward                 (unet.py:91-106)
as the reference actually runs it at inference (predict.py:25-35,118-123):
BatchNorm3d in TRAINING mode (batch statistics, biased variance, eps 1e-5),
max-pools with -inf padding (unet.py:166-187), depthwise transposed
convolutions (unet.py:216-242), crops and [upsampled, skip] concatenation
(unet.py:329-345), sigmoid head (unet.py:208-210).

Pinned: bit-identical to the verbatim reference module on the golden chunk
(tests/golden/unet_small.npz, tests/test_oracle_vs_reference.py).
"""
import numpy as np
import torch
import torch.nn.functional as F

ENCODER = [('c0', 1, 32), ('c1', 32, 64), ('c2', 64, 128), ('c3', 128, 256),
           ('c4', 256, 256)]
DECODER = [('c5_0', 512, 128), ('c6_0', 256, 64), ('c7_0', 128, 32), ('c8_0', 64, 5)]
UPS = [('up0', 256, (2, 2, 2)), ('up1', 128, (1, 2, 2)), ('up2', 64, (1, 2, 2)),
       ('up3', 32, (1, 2, 2))]
POOLS = [(1, 2, 2), (1, 2, 2), (1, 2, 2), (2, 2, 2)]   # all with padding (0,1,1)


def synth_state_dict(seed=0, out_channels=5):
    """Deterministic (numpy default_rng) state_dict in the reference's file
    format (train.py:414-420; key order of SURVEY.md Appendix A).  The bundled
    network file is missing from the reference checkout, so weights are
    synthesised: conv weights ~ U(-b, b), b = 1/sqrt(fan_in) (torch's default
    scale), BN weight ~ U(0.5, 1.5), BN bias ~ U(-0.3, 0.3)."""
    rng = np.random.default_rng(seed)
    sd = {}

    def conv(prefix, cin, cout, k=(3, 3, 3)):
        b = 1.0 / np.sqrt(cin * int(np.prod(k)))
        sd[prefix + '.weight'] = torch.from_numpy(
            rng.uniform(-b, b, (cout, cin) + tuple(k)).astype(np.float32))
        sd[prefix + '.bias'] = torch.from_numpy(rng.uniform(-b, b, cout).astype(np.float32))

    def bn(prefix, c):
        sd[prefix + '.weight'] = torch.from_numpy(rng.uniform(0.5, 1.5, c).astype(np.float32))
        sd[prefix + '.bias'] = torch.from_numpy(rng.uniform(-0.3, 0.3, c).astype(np.float32))
        sd[prefix + '.running_mean'] = torch.zeros(c)
        sd[prefix + '.running_var'] = torch.ones(c)
        sd[prefix + '.num_batches_tracked'] = torch.tensor(0, dtype=torch.long)

    mods = ENCODER + [(n, i, o if n != 'c8_0' else out_channels) for n, i, o in DECODER]
    for name, cin, cout in mods:
        conv(name + '.conv0', cin, cout)
        conv(name + '.conv1', cout, cout)
        bn(name + '.batch0', cout)
        bn(name + '.batch1', cout)
    for name, c, k in UPS:
        b = 1.0 / np.sqrt(int(np.prod(k)))
        sd[name + '.weight'] = torch.from_numpy(
            rng.uniform(-b, b, (c, 1) + k).astype(np.float32))
        sd[name + '.bias'] = torch.from_numpy(rng.uniform(-b, b, c).astype(np.float32))
    return sd


def _conv_module(x, sd, name, final, hook=None):
    for i in (0, 1):
        x = F.conv3d(x, sd[f'{name}.conv{i}.weight'], sd[f'{name}.conv{i}.bias'], padding=1)
        if hook is not None:
            hook(f'{name}.conv{i}', x)
        x = F.batch_norm(x, None, None, sd[f'{name}.batch{i}.weight'],
                         sd[f'{name}.batch{i}.bias'], training=True, momentum=0.1, eps=1e-5)
        if i == 0 or final == 'relu':
            x = F.relu(x)
        elif final == 'sigmoid':
            x = torch.sigmoid(x)
        if hook is not None:
            hook(f'{name}.act{i}', x)
    return x


def _up(x, sd, name, k):
    return F.conv_transpose3d(x, sd[name + '.weight'], sd[name + '.bias'], stride=k,
                              groups=x.shape[1])


@torch.no_grad()
def unet_forward(x, sd, hook=None):
    """x: (N,1,D,H,W) float32 -> (N,5,D,H,W) float32.  N>1 is NOT what the
    reference does (it runs batch 1, so statistics are per chunk): callers that
    batch must loop; this function asserts N == 1."""
    assert x.shape[0] == 1 and x.shape[1] == 1
    skips = []
    for (name, _, _), pool in zip(ENCODER[:-1], POOLS):
        x = _conv_module(x, sd, name, 'relu', hook)
        skips.append(x)
        x = F.max_pool3d(x, pool, stride=pool, padding=(0, 1, 1))
    x = _conv_module(x, sd, 'c4', 'relu', hook)
    crops = [(slice(None, -1),) * 2] * 3 + [(slice(1, -1),) * 2]
    for (up, _, k), (name, _, _), crop, skip in zip(UPS, DECODER, crops, skips[::-1]):
        x = _up(x, sd, up, k)
        x = x[(slice(None),) * 3 + crop]
        x = torch.cat([x, skip], 1)
        x = _conv_module(x, sd, name, 'sigmoid' if name == 'c8_0' else 'relu', hook)
    return x


def predict_frame(vol, sd, chunk_size=(10, 256, 256), margin=(1, 64, 64), threads=None):
    """Chunked prediction of one frame: (Z,Y,X) f32 -> (5,Z,Y,X) f32
    (predict.py:64-126 with the U-Net forced to the CPU)."""
    from .chunks import process_chunks
    if threads:
        torch.set_num_threads(threads)
    out = np.zeros((5,) + vol.shape, dtype=np.float32)

    def fn(chunk):
        t = torch.from_numpy(np.ascontiguousarray(chunk)[None, None])
        return unet_forward(t, sd)[0].numpy()

    return process_chunks(vol, chunk_size, out, margin, fn)
