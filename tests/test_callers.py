"""CPU checks of the c5 caller: LR schedule against its closed form, and the wiring of one
training step (encode -> normalise -> ladder -> backward -> AdamW -> scheduler) with tiny towers and the dense test
double for the kernels, single process and 2 gloo ranks under DistributedDataParallel."""
import math
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import sparsify_clip_b200 as scb
from callers.encoders import MiniCLIP
from callers.train_step import TrainStep, cosine_schedule_with_warmup
from sparsify_clip_b200 import backend_cuda
from tests._fake_backend import FakeBackend

TINY = dict(vision=dict(image_size=32, patch=16, width=32, layers=2, heads=2, out_dim=16),
            text=dict(vocab=50, context=8, width=32, layers=2, heads=2, out_dim=16))
CFG = {"loss_type": "only_lunif_n_then_anchor+lalign+lunif(centroids)", "only_lunif_epochs": 1, "anchor_temperature": 0.1,
       "anchor_temperature_learnable": True, "learning_rate": 1e-3, "fp16": False}


def test_lr_schedule_closed_form():
    p = torch.nn.Parameter(torch.zeros(1))
    for hold in (0, 1):
        opt = torch.optim.AdamW([p], lr=1.0)
        sch = cosine_schedule_with_warmup(opt, 20, 100, steps_sparsify=10, config={"only_lunif_epochs": hold})
        for step in range(100):
            lr = sch.get_last_lr()[0]
            if hold and step < 10:
                want = 1.0
            elif step < 20:
                want = step / 20.0
            else:
                want = max(0.0, 0.5 * (1.0 + math.cos(math.pi * (step - 20) / 80.0)))
            assert lr == pytest.approx(want, abs=1e-12), (hold, step)
            opt.step()
            sch.step()


def _data(B, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 3, 32, 32, generator=g), torch.randint(0, 50, (B, 8), generator=g)


def test_one_step_single_process():
    prev = backend_cuda.set_backend(FakeBackend())
    try:
        torch.manual_seed(0)
        model = MiniCLIP(**TINY)
        ts = TrainStep(model, CFG, t_total=10, steps_sparsify=2)
        before = [p.detach().clone() for p in model.parameters()]
        images, tokens = _data(6, 1)
        l0 = ts(images, tokens, epoch=0)          # L_unif-only warm-up epoch: (U_i + U_t) / 2
        l1 = ts(images, tokens, epoch=1)          # exp-4 composition
        assert torch.isfinite(l0) and torch.isfinite(l1) and ts.current_batch == 2
        assert any(not torch.equal(a, b) for a, b in zip(before, model.parameters()))
        assert ts.temperature.grad is not None and ts.temperature.item() != pytest.approx(0.1, abs=1e-9)
        # the first step matches the reference arithmetic written out in plain torch
        torch.manual_seed(0)
        ref = MiniCLIP(**TINY)
        i = ref.encode_image(images)
        t = ref.encode_text(tokens)
        i, t = i / i.norm(dim=-1, keepdim=True), t / t.norm(dim=-1, keepdim=True)
        want = 0.5 * (torch.pdist(i).pow(2).mul(-2).exp().mean().log() + torch.pdist(t).pow(2).mul(-2).exp().mean().log())
        assert l0.item() == pytest.approx(want.item(), rel=1e-5)
    finally:
        backend_cuda.set_backend(prev)


def _ddp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    backend_cuda.set_backend(FakeBackend())
    torch.manual_seed(0)
    model = torch.nn.parallel.DistributedDataParallel(MiniCLIP(**TINY))
    ts = TrainStep(model, CFG, t_total=10, group=dist.group.WORLD, steps_sparsify=0)
    images, tokens = _data(8, 1)
    n = 8 // world
    loss = ts(images[rank * n:(rank + 1) * n], tokens[rank * n:(rank + 1) * n], epoch=1)
    out[rank] = (loss.item(), [p.grad.detach().clone() for p in model.module.parameters()], ts.temperature.grad.item())
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_step_equals_full_batch_step():
    """2 ranks x 4 samples under DDP (loss scaled by the world size) == 1 process x 8 samples: the parameter gradients
    the optimiser sees (AdamW's first update is scale-invariant, so the gradients themselves are compared)."""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_ddp_worker, args=(2, 31517 + os.getpid() % 1000, out), nprocs=2, join=True)
    prev = backend_cuda.set_backend(FakeBackend())
    try:
        torch.manual_seed(0)
        model = MiniCLIP(**TINY)
        ts = TrainStep(model, CFG, t_total=10, steps_sparsify=0)
        images, tokens = _data(8, 1)
        loss = ts(images, tokens, epoch=1)
    finally:
        backend_cuda.set_backend(prev)
    for r in (0, 1):
        l, grads, dtau = out[r]
        assert l == pytest.approx(loss.item(), rel=1e-5)
        assert dtau == pytest.approx(ts.temperature.grad.item(), rel=1e-4)
        num = sum((a - b.grad).pow(2).sum().item() for a, b in zip(grads, model.parameters()))
        den = sum(b.grad.pow(2).sum().item() for b in model.parameters())
        assert math.sqrt(num / den) <= 1e-4, (r, math.sqrt(num / den))


@pytest.mark.gpu
@pytest.mark.parametrize("amp", [None, torch.bfloat16, torch.float16])
def test_train_step_on_gpu_matches_eager_reference_arithmetic(amp):
    """The c5 caller on the real library: first-step loss of both experiment-6 phases against the reference's op
    sequence in eager PyTorch on the same towers, and a parameter update that stays finite."""
    import torch.nn.functional as F
    cfg = dict(CFG, fp16=False)
    images, tokens = _data(96, 5)
    images, tokens = images.cuda(), tokens.cuda()
    for epoch in (0, 1):
        torch.manual_seed(0)
        model = MiniCLIP(**TINY).cuda()
        ts = TrainStep(model, cfg, t_total=10, amp_dtype=amp, steps_sparsify=0)
        with torch.no_grad(), torch.autocast("cuda", dtype=amp, enabled=amp is not None):
            i, t = model.encode_image(images).float(), model.encode_text(tokens).float()
        i, t = F.normalize(i, dim=-1), F.normalize(t, dim=-1)
        lunif = lambda x: torch.pdist(x).pow(2).mul(-2).exp().mean().log()
        if epoch == 0:
            want = 0.5 * (lunif(i) + lunif(t))
        else:
            logits = i @ t.t() / 0.1
            tgt = torch.arange(96, device="cuda")
            want = (0.5 * (F.cross_entropy(logits, tgt) + F.cross_entropy(logits.t(), tgt)) + (i - t).norm(dim=1).pow(2).mean()
                    + lunif(F.normalize((i + t) / 2, dim=-1)))
        got = ts(images, tokens, epoch=epoch)
        tol = 1e-4 if amp is None else 3e-2          # autocast rounds the embeddings to 8 / 11 bits before the loss
        assert abs(got.item() - want.item()) <= tol * max(1.0, abs(want.item())), (amp, epoch, got.item(), want.item())
        assert all(torch.isfinite(p).all() for p in model.parameters())
        assert ts.temperature.grad is None or torch.isfinite(ts.temperature.grad).all()


@pytest.mark.gpu
@pytest.mark.parametrize("epoch", [0, 1])
def test_train_step_parameter_gradients_match_the_eager_reference_chain(epoch):
    """One c5 step, fp32: the gradient that reaches EVERY encoder parameter (through pre-loss normalise -> loss ladder,
    sparsify_clip.py:768-938) and the learnable temperature, against the reference's op sequence in eager PyTorch
    autograd on an identical copy of the towers."""
    import copy

    import torch.nn.functional as F
    cfg = dict(CFG, fp16=False, anchor_temperature_learnable=True)
    images, tokens = _data(96, 7)
    images, tokens = images.cuda(), tokens.cuda()
    torch.manual_seed(0)
    model = MiniCLIP(**TINY).cuda()
    ref_model = copy.deepcopy(model)
    ts = TrainStep(model, cfg, t_total=10, amp_dtype=None, steps_sparsify=0)
    img, txt = ts.embed(images, tokens)
    loss = ts.loss(img, txt, epoch)
    (loss * 7.0).backward()
    # the reference's own lines: :768-773 encode + normalise, :796-809 the exp-6 ladder
    tau = torch.nn.Parameter(torch.tensor(0.1))
    i, t = ref_model.encode_image(images), ref_model.encode_text(tokens)
    i = i / i.norm(dim=-1, keepdim=True)
    t = t / t.norm(dim=-1, keepdim=True)
    lunif = lambda x: torch.pdist(x, p=2).pow(2).mul(-2).exp().mean().log()
    if epoch == 0:
        want = (lunif(i) + lunif(t)) / 2
    else:
        logits = i @ t.t() / tau
        tgt = torch.arange(96, device="cuda")
        want = ((F.cross_entropy(logits, tgt) + F.cross_entropy(logits.t(), tgt)) / 2 + (i - t).norm(dim=1).pow(2).mean()
                + lunif(F.normalize((i + t) / 2, dim=-1)))
    (want * 7.0).backward()
    assert abs(loss.item() - want.item()) <= 1e-5 * max(1.0, abs(want.item()))
    num = den = 0.0
    for (name, p), q in zip(model.named_parameters(), ref_model.parameters()):
        if q.grad is None:
            assert p.grad is None or p.grad.abs().max().item() == 0.0, name
            continue
        num += (p.grad.double() - q.grad.double()).pow(2).sum().item()
        den += q.grad.double().pow(2).sum().item()
    assert math.sqrt(num / den) <= 2e-4, math.sqrt(num / den)      # fp32 encoders: the two chains differ by fp32 rounding only
    if epoch == 1:
        assert ts.temperature.grad.item() == pytest.approx(tau.grad.item(), rel=1e-4)
