"""One optimisation step of the reference training loop (sparsify_clip.py:750-965) with the loss path replaced by
sparsify_clip_b200: encode -> [pre-loss normalise (:772-773) -> loss_type ladder (:775-938)] as ONE node -> backward ->
AdamW (:730) -> LR schedule (:68-107, :735-737).

Sharded use (one process per GPU): the encoders are wrapped in DistributedDataParallel by the caller, `group` is
passed through to compose_loss, which returns the full-batch loss and this rank's slice of the full-batch gradient.
DDP AVERAGES parameter gradients over ranks, so the loss is multiplied by the world size before backward: the
averaged gradient is then exactly the gradient of the full-batch loss (the semantics of the reference's
single-process DataParallel run)."""
import math

import torch
import torch.distributed as dist
from torch.optim.lr_scheduler import LambdaLR

import sparsify_clip_b200 as scb


def cosine_schedule_with_warmup(optimizer, num_warmup_steps, num_training_steps, num_cycles=0.5, last_epoch=-1,
                                steps_sparsify=462, config=None):
    """sparsify_clip.py:66-107: constant LR while the L_unif-only warm-up epoch runs (only_lunif_epochs > 0), linear
    warm-up to the optimiser's LR, then half a cosine to zero."""
    hold = bool(config and config.get("only_lunif_epochs", 0) > 0)

    def lr_lambda(step):
        if hold and step < steps_sparsify:
            return 1.0
        if step < num_warmup_steps:
            return float(step) / float(max(1, num_warmup_steps))
        progress = float(step - num_warmup_steps) / float(max(1, num_training_steps - num_warmup_steps))
        return max(0.0, 0.5 * (1.0 + math.cos(math.pi * float(num_cycles) * 2.0 * progress)))

    return LambdaLR(optimizer, lr_lambda, last_epoch)


class TrainStep:
    """`model` exposes encode_image / encode_text (or is a DDP wrapper whose forward returns both embeddings).
    config: the reference's YAML keys (loss_type, only_lunif_epochs, anchor_temperature, anchor_temperature_learnable,
    learning_rate, fp16, the beta/alpha schedule keys)."""

    def __init__(self, model, config, t_total, *, group=None, amp_dtype=None, steps_sparsify=462, fuse_normalize=True):
        self.model, self.config, self.t_total, self.group = model, config, int(t_total), group
        # the pre-loss normalise (:772-773) inside the loss node: its backward rides on the node's gradient combine pass and
        # the encoders receive d loss / d (raw embedding) directly; False = a separate l2_normalize node in front
        self.fuse_normalize = bool(fuse_normalize)
        self.world = dist.get_world_size(group) if group is not None else 1
        self.temperature = config["anchor_temperature"]
        params = list(model.parameters())
        if config.get("anchor_temperature_learnable"):
            # a CPU nn.Parameter, as in the reference (:717): its gradient comes back on its own device
            self.temperature = torch.nn.Parameter(torch.tensor(float(self.temperature), dtype=torch.float32))
            params.append(self.temperature)
        self.optimizer = torch.optim.AdamW(params, lr=config["learning_rate"])
        # the reference autocasts to fp16 with a GradScaler (:731, :961); bf16 needs no scaler
        self.amp_dtype = amp_dtype if amp_dtype is not None else (torch.float16 if config.get("fp16") else None)
        dev_type = next(model.parameters()).device.type
        self.dev_type = dev_type
        self.scaler = torch.amp.GradScaler(dev_type) if self.amp_dtype == torch.float16 and dev_type == "cuda" else None
        self.scheduler = cosine_schedule_with_warmup(self.optimizer, int(0.20 * self.t_total), self.t_total,
                                                     steps_sparsify=steps_sparsify, config=config)
        self.current_batch = 0          # never reset across epochs (:750, :755)

    def embed(self, images, tokens):
        m = self.model
        if hasattr(m, "encode_image"):
            img, txt = m.encode_image(images), m.encode_text(tokens)
        else:                           # DDP wrapper: one forward so that its reducer sees one backward
            img, txt = m(images, tokens)
        if self.fuse_normalize:
            return img, txt                 # raw encoder outputs: loss() normalises them
        return scb.l2_normalize(img), scb.l2_normalize(txt)

    def loss(self, img, txt, epoch):
        return scb.compose_loss(self.config, img, txt, self.temperature, epoch=epoch, current_batch=self.current_batch,
                                t_total=self.t_total, group=self.group, normalize=self.fuse_normalize)

    def __call__(self, images, tokens, epoch=0):
        self.current_batch += 1
        with torch.autocast(device_type=self.dev_type, dtype=self.amp_dtype, enabled=self.amp_dtype is not None):
            img, txt = self.embed(images, tokens)
            loss = self.loss(img, txt, epoch)
        self.optimizer.zero_grad()
        back = loss * float(self.world) if self.world > 1 else loss
        if self.scaler is not None:
            self.scaler.scale(back).backward()
            self._sync_temperature_grad()
            self.scaler.step(self.optimizer)
            self.scaler.update()
        else:
            back.backward()
            self._sync_temperature_grad()
            self.optimizer.step()
        self.scheduler.step()
        return loss.detach()

    def _sync_temperature_grad(self):
        # the temperature is outside the DDP-wrapped module: undo the world-size factor (every rank already holds the
        # full-batch d/dtau)
        if self.world > 1 and isinstance(self.temperature, torch.nn.Parameter) and self.temperature.grad is not None:
            self.temperature.grad.div_(float(self.world))
