"""project_to_mel on the tcgen05 tensor cores (csrc/project.cu) against torch: the reference computes
``self.project_to_mel(src_emb)`` (model.py:249) with ``nn.Linear`` under bf16 autocast.

Tolerance: both sides round inputs, weight and bias to bf16, accumulate in float32 and round the result to bf16; they
differ in the order of the float32 accumulation only.  Against the exact (float64) product of the bf16-rounded operands
a correctly rounded bf16 result is within half a bf16 ulp (2^-9 relative); the test allows one ulp (2^-8) plus the
float32 accumulation error, and requires agreement with torch's own autocast Linear to one bf16 ulp."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _exact(x, w, b):
    xb, wb = x.to(torch.bfloat16).double(), w.to(torch.bfloat16).double()
    y = xb @ wb.t()
    return y if b is None else y + b.to(torch.bfloat16).double()


@pytest.mark.parametrize("rows,n_out,bias", [(64 * 246, 768, True), (1, 768, True), (129, 256, False), (1000, 32, True),
                                            (128 * 149 + 7, 512, True)])
def test_projection_matches_autocast_linear(rows, n_out, bias):
    from adt_str_b200 import ProjectToMel
    torch.manual_seed(rows + n_out)
    dev = torch.device("cuda", 0)
    lin = torch.nn.Linear(128, n_out, bias=bias).to(dev)
    proj = ProjectToMel.from_linear(lin)
    x = torch.rand(rows, 128, device=dev)                 # log-mel values lie in [0, 1]
    with torch.no_grad():
        got = proj(x)
        with torch.autocast("cuda", torch.bfloat16):
            want = lin(x)
    assert got.dtype == torch.bfloat16 and got.shape == want.shape == (rows, n_out)
    exact = _exact(x, lin.weight, lin.bias)
    ulp = exact.abs() * 2.0 ** -8 + 1e-6
    assert bool(((got.double() - exact).abs() <= ulp).all())
    assert bool(((want.double() - exact).abs() <= ulp).all())              # the library is within the same bound
    assert bool(((got.double() - want.double()).abs() <= exact.abs() * 2.0 ** -7 + 1e-6).all())
    frac_equal = float((got == want).float().mean())
    assert frac_equal > 0.98, frac_equal                                   # and bit-equal almost everywhere


@pytest.mark.parametrize("rows,n_out,bias", [(128 * 700 + 5, 768, True), (128 * 40, 512, True), (128 * 9 + 127, 256, False),
                                            (300, 32, True), (128 * 3, 416, True), (128 * 300, 96, True),
                                            (128 * 151 + 33, 672, True)])
def test_every_part_and_chunk_width(rows, n_out, bias):
    """The column parts (<= 384 columns per CTA), the 128-column accumulator chunks and the drain's two store paths (a
    TMA tensor store for a warp's 64 columns, a staged store for a lone 32-column unit) over output widths that mix them
    (416 -> 224 + 192 -> chunks 128 + 96 / 128 + 64; 96 -> one chunk of three units; 32 -> half the drain warps idle),
    over many tiles per CTA (the accumulator ring and the A buffers wrap) and a ragged last tile (the TMA clips it).
    Input values outside [0, 1] and of both signs."""
    from adt_str_b200 import ProjectToMel
    torch.manual_seed(rows * 3 + n_out)
    dev = torch.device("cuda", 0)
    lin = torch.nn.Linear(128, n_out, bias=bias).to(dev)
    proj = ProjectToMel.from_linear(lin).eval()
    x = torch.rand(rows, 128, device=dev) * 4 - 2
    with torch.no_grad():
        got = proj(x)
        again = proj(x)
        with torch.autocast("cuda", torch.bfloat16):
            want = lin(x)
    torch.cuda.synchronize()
    assert torch.equal(got.view(torch.int16), again.view(torch.int16))      # deterministic
    exact = _exact(x, lin.weight, lin.bias)
    assert bool(((got.double() - exact).abs() <= exact.abs() * 2.0 ** -8 + 1e-6).all())
    assert float((got == want).float().mean()) > 0.98


def test_projection_writes_nothing_beyond_its_rows():
    """Rows beyond n_rows of the last tile are clipped by the tensor map (and by the staged store's row test): the
    memory behind the output stays untouched."""
    from adt_str_b200 import ProjectToMel, _lib
    dev = torch.device("cuda", 0)
    torch.manual_seed(3)
    for n_out in (768, 416):
        proj = ProjectToMel(128, n_out).to(dev).eval()
        rows = 128 * 2 + 37
        buf = torch.full((rows + 200, n_out), 5.0, dtype=torch.bfloat16, device=dev)
        x = torch.rand(rows, 128, device=dev)
        native = proj._handle(dev)
        _lib.check(native.lib.adtfe_linear_forward(native.handle, x.data_ptr(), rows, buf.data_ptr(),
                                                   torch.cuda.current_stream(dev).cuda_stream), "adtfe_linear_forward")
        torch.cuda.synchronize()
        with torch.no_grad():
            assert torch.equal(buf[:rows], proj(x))
        assert bool((buf[rows:] == 5.0).all())


def test_projection_shapes_training_and_errors():
    from adt_str_b200 import ProjectToMel, _lib
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    proj = ProjectToMel(128, 768).to(dev)
    assert sorted(k for k, _ in proj.named_parameters()) == ["bias", "weight"]
    x = torch.rand(4, 246, 128, device=dev)
    y = proj(x)                                            # training mode: through the autograd function
    assert y.shape == (4, 246, 768) and y.dtype == torch.bfloat16 and y.requires_grad
    y.float().pow(2).sum().backward()
    ref = torch.nn.Linear(128, 768).to(dev)
    ref.load_state_dict(proj.state_dict())
    with torch.autocast("cuda", torch.bfloat16):
        ref(x).float().pow(2).sum().backward()
    assert torch.allclose(proj.weight.grad, ref.weight.grad, rtol=2e-2, atol=2e-2 * float(ref.weight.grad.abs().max()))
    assert torch.allclose(proj.bias.grad, ref.bias.grad, rtol=2e-2, atol=2e-2 * float(ref.bias.grad.abs().max()))
    # an optimiser step changes the weights: the device image follows
    with torch.no_grad():
        before = proj(x).clone()
        proj.weight.mul_(0.5)
        after = proj(x)
    assert not torch.equal(before, after)
    with pytest.raises(_lib.AdtfeError):
        ProjectToMel(64, 768).to(dev)(torch.rand(2, 64, device=dev))
    with pytest.raises(RuntimeError, match="no CPU path"):
        ProjectToMel(128, 768)(torch.rand(2, 128))


@pytest.mark.timeout(120)
def test_projections_on_concurrent_streams_share_the_tensor_memory():
    """Every CTA of the kernel allocates all 512 columns of its SM's tensor memory.  Two launches on two streams, and a
    library bf16 GEMM (tcgen05 too) on a third, must neither deadlock on tcgen05.alloc nor disturb each other's
    accumulators: the results equal the ones computed alone."""
    from adt_str_b200 import ProjectToMel
    dev = torch.device("cuda", 0)
    torch.manual_seed(11)
    p1, p2 = ProjectToMel(128, 768).to(dev).eval(), ProjectToMel(128, 512).to(dev).eval()
    x1, x2 = torch.rand(128 * 600 + 3, 128, device=dev), torch.rand(128 * 450, 128, device=dev)
    a, b = torch.randn(4096, 4096, device=dev, dtype=torch.bfloat16), torch.randn(4096, 4096, device=dev, dtype=torch.bfloat16)
    with torch.no_grad():
        want1, want2, want3 = p1(x1).clone(), p2(x2).clone(), (a @ b).clone()
        torch.cuda.synchronize()
        s1, s2, s3 = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        for _ in range(5):
            with torch.cuda.stream(s1):
                got1 = p1(x1)
            with torch.cuda.stream(s3):
                got3 = a @ b
            with torch.cuda.stream(s2):
                got2 = p2(x2)
        torch.cuda.synchronize()
    assert torch.equal(got1, want1) and torch.equal(got2, want2) and torch.equal(got3, want3)


def test_projection_of_the_logmel_output():
    """The log-mel matrix adtfe_logmel writes feeds the projection directly (rows = B * T, 128 floats)."""
    from adt_str_b200 import ComputeMelSpectrogram, ProjectToMel
    dev = torch.device("cuda", 0)
    mel = ComputeMelSpectrogram(24000, 2048, 0.01, 128)
    proj = ProjectToMel(128, 768).to(dev).eval()
    wave = torch.randn(8, 61440, generator=torch.Generator().manual_seed(1)).to(dev)
    with torch.no_grad():
        feat = mel(wave)
        emb = proj(feat)
        with torch.autocast("cuda", torch.bfloat16):
            want = torch.nn.functional.linear(feat, proj.weight, proj.bias)
    assert emb.shape == (8, 246, 768)
    assert float((emb.float() - want.float()).abs().max()) <= 2.0 ** -7 * float(want.float().abs().max()) + 1e-6
