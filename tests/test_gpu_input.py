"""GPU input pipeline (pggan_b200/data.py, csrc/input.cu) against the reference's transform chain
(lib/dataset.py:100-111) evaluated by torchvision itself on float tensors, image by image, under the same seed."""
import pytest
import torch
import torchvision.transforms as T
import torchvision.transforms.functional as TF

pytestmark = pytest.mark.gpu


def _reference_chain(size):
    """the reference's Compose with ToTensor replaced by its tensor equivalent (the chain runs on a float tensor)"""
    return T.Compose([
        T.Resize((size, size), antialias=True),
        T.RandomHorizontalFlip(p=0.5),
        T.ColorJitter(0.2, 0.2, 0.2, 0.01),
        T.Normalize((0.5, 0.5, 0.5), (0.5, 0.5, 0.5)),
    ])


@pytest.mark.parametrize("scale_index,src", [(0, (37, 53)), (1, (64, 64)), (3, (218, 178)), (4, (64, 96)), (6, (300, 260))])
def test_pipeline_matches_torchvision(scale_index, src):
    from pggan_b200 import data
    size = 2 ** (scale_index + 2)
    B = 6
    g = torch.Generator().manual_seed(3 + scale_index)
    imgs = torch.randint(0, 256, (B, 3) + src, dtype=torch.uint8, generator=g)
    imgs[0] = 128                                  # a flat grey image: the hue path's maxc == minc branch
    imgs[1, :, : src[0] // 2] //= 8                # a dark half
    chain = _reference_chain(size)
    ref = []
    for b in range(B):                             # the reference transforms one image at a time (Dataset.__getitem__)
        torch.manual_seed(1000 + b)
        ref.append(chain(imgs[b].float() / 255.0))
    ref = torch.stack(ref)
    params_i, params_f = [], []
    for b in range(B):                             # the same draws, in the same order, from the same seeds
        torch.manual_seed(1000 + b)
        ip, fp = data.sample_params(1)
        params_i.append(ip)
        params_f.append(fp)
    ip, fp = torch.cat(params_i), torch.cat(params_f)
    assert ip[:, 0].sum() not in (0, B) or B < 4   # both flipped and unflipped samples present
    pipe = data.GPUInputPipeline(scale_index)
    out = pipe(imgs.cuda(), params=(ip, fp)).cpu()
    assert out.shape == (B, 3, size, size) and out.dtype == torch.float32
    err = (out - ref).abs()
    # fp32 reassociation only — except where a pixel sits on a branch of the piecewise hue conversion (floor(6h), the
    # max channel): there an ulp decides the branch and both branches agree to first order
    assert float(err.max()) < 2e-3, float(err.max())
    assert float((err > 2e-5).float().mean()) < 2e-3, float((err > 2e-5).float().mean())
    # resize alone: against torchvision / ATen's anti-aliased bilinear kernel
    r = data.resize_antialias(imgs.cuda(), size).cpu()
    r_ref = TF.resize(imgs.float() / 255.0, [size, size], antialias=True)
    assert float((r - r_ref).abs().max()) < 2e-6
    # flip and normalisation are exact: no jitter (identity factors), all flipped
    ip2 = torch.tensor([[1, 0, 1, 2, 3]] * B, dtype=torch.int32)
    fp2 = torch.tensor([[1.0, 1.0, 1.0, 0.0]] * B)
    out2 = data.jitter_normalize_fade(r.cuda(), ip2, fp2).cpu()
    hue_id = torch.stack([TF.adjust_hue(x, 0.0) for x in r_ref])          # hue 0 still round-trips through HSV
    assert float((out2 - (hue_id.flip(-1) - 0.5) / 0.5).abs().max()) < 1e-5


def test_pipeline_fade_equals_train_step_fade():
    """fade fused into the pipeline == ops.ImageFade on the un-faded pipeline output (pggan/model.py:217-221), with
    alpha as a float and as the device scalar a captured graph reads"""
    from pggan_b200 import data, ops
    torch.manual_seed(0)
    imgs = torch.randint(0, 256, (4, 3, 40, 40), dtype=torch.uint8).cuda()
    ip, fp = data.sample_params(4)
    pipe = data.GPUInputPipeline(2)
    plain = pipe(imgs, params=(ip, fp))
    faded = pipe(imgs, fade_alpha=0.3, params=(ip, fp))
    ref = ops.ImageFade.apply(plain, None, 0.3)
    assert float((faded - ref).abs().max()) < 1e-6
    adev = torch.full((1,), 0.3, device="cuda")
    assert torch.equal(pipe(imgs, fade_alpha=ops.AlphaRef(0.3, adev), params=(ip, fp)), faded)
    assert float(plain.min()) >= -1.0 and float(plain.max()) <= 1.0
