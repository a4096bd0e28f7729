"""Data-parallel GAN training around the native generator (BASELINE configs[2]/[3]).

The reference trains on one GPU (``Trainer._train_step``, harana/bin/train_fastsvc.py:157-235).  Here one process per
GPU runs the same step on its own batch of 16 and the replicas meet only in the gradients:

  * ``GradBucket`` keeps every gradient of a model in ONE flat fp32 buffer (``p.grad`` are views into it), so a step
    needs one NCCL all-reduce per model -- generator 2.75 M floats (11 MB), discriminator 4.35 M (MelGAN MSD) or
    70.7 M (HiFiGAN MSMPD, 283 MB) -- over NVLink / NVSwitch, and the global-norm clip of the reference
    (``clip_grad_norm_`` at :201-205, :229-233) is a single norm of that buffer taken AFTER the reduce, so it is the
    norm of the global gradient.
  * ``GanTrainer.step`` is ``_train_step`` with those two reduces.  The discriminator's all-reduce is asynchronous
    and is completed (wait, clip, optimizer step) only where the discriminator is next needed -- the adversarial term
    of the next generator phase -- so it overlaps the next generator forward and STFT loss.

Inference needs none of this (no collective on the forward path).  Losses, discriminator and optimizers are whatever
host PyTorch objects the caller passes (north_star keeps them in PyTorch).
"""

import torch
import torch.distributed as dist


class GradBucket:
    """All gradients of ``params`` in one flat fp32 buffer; ``p.grad`` is a view into it."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("GradBucket needs at least one parameter that requires grad")
        dev = self.params[0].device
        if any(p.device != dev or p.dtype != torch.float32 for p in self.params):
            raise ValueError("GradBucket: all parameters must be fp32 on one device")
        self.group = group
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n
        self._work = None

    @property
    def world(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def zero(self):
        """Replaces ``optimizer.zero_grad()`` (which would drop the views with set_to_none=True)."""
        self.wait()
        self.flat.zero_()

    def check_views(self):
        off = 0
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + 4 * off:
                raise RuntimeError("GradBucket: a parameter's .grad no longer aliases the bucket (zero_grad(set_to_none="
                                   "True) or a .to() call?); use bucket.zero() and rebuild the bucket after moving")
            off += p.numel()

    def all_reduce(self, async_op=False):
        """Average the bucket over the data-parallel group (sum -> mean, like DDP)."""
        self.check_views()
        if self.world == 1:
            return
        self.flat.div_(self.world)
        self._work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        if not async_op:
            self.wait()

    def wait(self):
        if self._work is not None:
            self._work.wait()      # on CUDA: the current stream waits for the NCCL stream
            self._work = None

    def clip_(self, max_norm, eps=1e-6):
        """``torch.nn.utils.clip_grad_norm_(params, max_norm)`` on the reduced gradient; returns the total norm."""
        self.wait()
        total = torch.linalg.vector_norm(self.flat, 2)
        if max_norm and max_norm > 0:
            self.flat.mul_(torch.clamp(max_norm / (total + eps), max=1.0))
        return total


class GanTrainer:
    """``Trainer._train_step`` (train_fastsvc.py:157-235) for one data-parallel replica."""

    def __init__(self, generator, discriminator, stft_loss, gen_adv_loss, dis_adv_loss, opt_g, opt_d,
                 lambda_adv=2.5, lambda_aux=1.0, generator_grad_norm=10.0, discriminator_grad_norm=1.0,
                 sched_g=None, sched_d=None, group=None):
        self.G, self.D = generator, discriminator
        self.stft_loss, self.gen_adv_loss, self.dis_adv_loss = stft_loss, gen_adv_loss, dis_adv_loss
        self.opt_g, self.opt_d, self.sched_g, self.sched_d = opt_g, opt_d, sched_g, sched_d
        self.lambda_adv, self.lambda_aux = lambda_adv, lambda_aux
        self.g_norm, self.d_norm = generator_grad_norm, discriminator_grad_norm
        self.gb = GradBucket(generator.parameters(), group)
        self.db = GradBucket(discriminator.parameters(), group) if discriminator is not None else None
        self._d_pending = False

    def finish_discriminator_step(self):
        """Complete the deferred discriminator update: wait for its all-reduce, clip, step (:229-235)."""
        if not self._d_pending:
            return
        self.db.clip_(self.d_norm)
        self.opt_d.step()
        if self.sched_d is not None:
            self.sched_d.step()
        self._d_pending = False

    def step(self, x, y, adversarial=True):
        """One training step on this replica's batch: x = (ppg, sine, lft[, spk]) and target y, already on the device.
        Returns the (detached) loss tensors; no host sync happens here."""
        logs = {}
        # ---- generator (:166-208) ----
        y_ = self.G(*x)
        sc_loss, mag_loss = self.stft_loss(y_, y)
        gen_loss = (sc_loss + mag_loss) * self.lambda_aux
        logs["spectral_convergence_loss"], logs["log_stft_magnitude_loss"] = sc_loss.detach(), mag_loss.detach()
        if adversarial:
            self.finish_discriminator_step()          # the critic must be up to date before it scores y_
            adv_loss = self.gen_adv_loss(self.D(y_))
            gen_loss = gen_loss + self.lambda_adv * adv_loss
            logs["adversarial_loss"] = adv_loss.detach()
        logs["generator_loss"] = gen_loss.detach()
        self.gb.zero()
        gen_loss.backward()
        self.gb.all_reduce()
        logs["generator_grad_norm"] = self.gb.clip_(self.g_norm)
        self.opt_g.step()
        if self.sched_g is not None:
            self.sched_g.step()
        # ---- discriminator (:212-235) ----
        if adversarial:
            with torch.no_grad():
                y_ = self.G(*x)                        # re-computed with the updated generator (:214-215)
            p = self.D(y)
            p_ = self.D(y_.detach())
            real_loss, fake_loss = self.dis_adv_loss(p_, p)
            dis_loss = real_loss + fake_loss
            logs["real_loss"], logs["fake_loss"] = real_loss.detach(), fake_loss.detach()
            self.db.zero()
            dis_loss.backward()
            self.db.all_reduce(async_op=True)          # completes in finish_discriminator_step()
            self._d_pending = True
        return logs
