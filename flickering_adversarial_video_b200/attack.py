"""The flickering-attack optimisation loop: one `FlickerAttack.step()` == one
`sess.run([train_op, loss, …])` of the reference (i3d_adversarial_main_single_video_npy.py:213-215,
i3d_adversarial_main_single_class_gen.py:245-247, i3d_adversarial_main_universal.py:163-175):
apply delta -> forward -> adversarial loss -> backward to delta -> (sum over ranks) -> regulariser
gradients + Adam.  1 forward + 1 backward-to-input per step; no weight gradients, no extra forwards.

Multi-GPU (universal / class-generalisation attacks): one process per GPU, the clip batch is sharded
on dim 0, delta / Adam state are replicated and the packed `[T,3] + scalars` buffer is sum-all-reduced
once per step (torch.distributed, NCCL over NVLink).  The margin loss is a SUM over samples
(utils/kinetics_i3d_utils.py:285) so the global gradient is the plain sum of the rank gradients; the
CE loss is a MEAN (:305) and each rank divides by the global batch.  Regulariser gradients depend on
delta only and are added once, after the all-reduce, identically on every rank.
"""
import torch

from . import _lib as L
from . import dist as fdist
from .engine import EvalEngine, FlickerEngine


class FlickerAttack:
    def __init__(self, weights, batch, frames, attack_cfg=None, height=None, width=None, num_classes=400,
                 device=0, lr=1e-3, stack=None, delta_clip=None, process_group=None, arch="i3d", sharded=True,
                 frame_range=None):
        """arch "i3d": TF-stack rules, delta_clip 0.4 (utils/kinetics_i3d_utils.py:104-105).
        arch "r3d_18"/"mc3_18"/"r2plus1d_18": torch-stack rules (model.py:58-250): delta_clip is
        l_inf_pert_norm, the regulariser is beta_1*thick + (1-beta_1)*(diff+lap) on the clamped delta.
        sharded=False: this attack is a per-rank replica (single-video attacks: every rank works on its own
        video) and never joins a collective even when torch.distributed is initialised.
        frame_range=(start, end): the reference's `_IND_START` / `_IND_END` frame mask (utils/kinetics_i3d_utils.py:14-15,
        107-113: delta acts on frames start..end inclusive, the out-of-range index `end == T` is dropped by one_hot);
        None = the reference's default, every frame."""
        self.eng = FlickerEngine(batch, frames, height, width, num_classes, device, arch=arch)
        self.eng.load_weights(weights)
        self._weights = weights      # the evaluation handle (evaluator()) packs its own forward weights from these
        self._eval = None
        self.arch = arch
        if stack is None:
            stack = "torch" if self.eng.torch_stack else "tf"
        if delta_clip is None:
            delta_clip = 0.1 if self.eng.torch_stack else 0.4   # r2plus1d_main_universal_attack.py:45
        self.device = self.eng.device
        self.B, self.T, self.K = batch, frames, num_classes
        self.stack = L.FAV_STACK_TF if stack == "tf" else L.FAV_STACK_TORCH
        self.lr = lr
        self.delta_clip = delta_clip
        cfg = dict(attack_cfg or {})
        self.improve_loss = bool(cfg.get("IMPROVE_ADV_LOSS", True))
        self.targeted = bool(cfg.get("TARGETED_ATTACK", False))
        self.use_logits = bool(cfg.get("USE_LOGITS", False))
        self.margin = float(cfg.get("PROB_MARGIN", 0.05))
        self.beta0 = float(cfg.get("LAMBDA", 1.0))
        self.beta1 = float(cfg.get("BETA_1", 0.5))
        self.beta2 = float(cfg.get("BETA_2", 0.5))
        self.beta3 = float(cfg.get("BETA_2", 0.5))   # the drivers set _beta_3 = BETA_2 (single_video_npy.py:98)
        if self.stack == L.FAV_STACK_TORCH:
            # Losses.flickering_regularization_loss (model.py:198-209): beta_1*norm + (1-beta_1)*(diff + lap)
            self.beta2 = self.beta3 = 1.0 - self.beta1
        # distributed
        self.pg = process_group
        self.world = 1
        if sharded and torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size(process_group)
        self.global_batch = batch * self.world
        # state: delta (eps_rgb, utils/kinetics_i3d_utils.py:100 — zeros), Adam slots, step
        n = frames * 3
        self.delta = torch.zeros((frames, 3), dtype=torch.float32, device=self.device)
        self.m = torch.zeros_like(self.delta)
        self.v = torch.zeros_like(self.delta)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=self.device)
        # packed communication buffer: [T*3 gradient | FAV_S_COUNT scalars]
        self.comm = torch.zeros(n + L.S_COUNT, dtype=torch.float32, device=self.device)
        self.grad = self.comm[:n].view(frames, 3)
        self.scalars = self.comm[n:]
        self.frame_mask = None
        if frame_range is not None and tuple(frame_range) != (0, frames) and tuple(frame_range) != (0, frames - 1):
            start, end = int(frame_range[0]), min(int(frame_range[1]), frames - 1)
            self.frame_mask = torch.zeros((frames, 1), dtype=torch.float32, device=self.device)
            self.frame_mask[start:end + 1] = 1.0
        self.eng.grad = self.grad
        self.eng.scalars = self.scalars
        # host staging for the end-to-end path
        self._stage = None
        self._copy_stream = None
        self._host_scalars = None

    # ---- state ---------------------------------------------------------------------------
    def reset(self):
        """sess.run(eps_rgb.initializer) + re-initialise the optimizer slots
        (single_video_npy.py:205-206)."""
        self.delta.zero_()
        self.m.zero_()
        self.v.zero_()
        self.step_count.zero_()

    @property
    def perturbation(self):
        """the reference's eps_rgb, shape [T,1,1,3]"""
        return self.delta.view(self.T, 1, 1, 3)

    def state_dict(self):
        return {"delta": self.delta.cpu(), "m": self.m.cpu(), "v": self.v.cpu(), "step": int(self.step_count.item())}

    def load_state_dict(self, sd):
        self.delta.copy_(sd["delta"])
        self.m.copy_(sd["m"])
        self.v.copy_(sd["v"])
        self.step_count.fill_(int(sd["step"]))

    # ---- one optimisation step -------------------------------------------------------------
    def step(self, clips, labels, adv_flag=1.0, lr=None):
        """clips: DEVICE uint8/float32 [B,T,H,W,3]; labels: DEVICE int64 [B] (the target class ids
        for a targeted attack).  Asynchronous; returns the device scalar block (see _lib.S_*)."""
        e = self.eng
        e.apply(clips, self._applied_delta(), adv_flag=adv_flag, delta_clip=self.delta_clip)
        e.forward()
        gscale = 1.0
        e.loss(labels, improve_loss=self.improve_loss, targeted=self.targeted, use_logits=self.use_logits,
               margin=self.margin, grad_scale=gscale, global_batch=self.global_batch, stack=self.stack)
        e.backward()
        if self.frame_mask is not None:
            self.grad.mul_(self.frame_mask)
        if self.world > 1:
            fdist.allreduce_sum_(self.comm, self.pg)
        e.update(self.delta, self.grad, self.m, self.v, self.step_count, self.beta0, self.beta1, self.beta2,
                 self.beta3, lr=self.lr if lr is None else lr, delta_clip=self.delta_clip, stack=self.stack)
        return self.scalars

    def check_replicas(self):
        """Raise if the replicated perturbation has diverged between ranks (sharded attacks only; a collective: every
        rank must call it at the same step)."""
        if self.world > 1 and not fdist.replicas_equal(self.delta, self.pg):
            raise RuntimeError("the perturbation differs between ranks: the replicas have diverged")

    def _applied_delta(self):
        """mask_rgb * eps_rgb (utils/kinetics_i3d_utils.py:128): delta as the network sees it"""
        return self.delta if self.frame_mask is None else self.delta * self.frame_mask

    def step_rolled(self, clips, labels, shift, adv_flag=1.0, lr=None):
        """Cyclic perturbation attack: the network sees roll(delta, shift) along T (TF: `tf.roll(input_pert,
        random_shift_2, axis=0)`, utils/kinetics_i3d_utils.py:130-131,137; torch: `torch.roll(perturbation_normalize,
        shifts, dims=1)`, model.py:91-92) and the gradient is rolled back onto delta.  The clamp mask of the update
        commutes with the roll, the regularisers are roll-invariant (circular differences)."""
        e = self.eng
        rolled = torch.roll(self._applied_delta(), int(shift), dims=0).contiguous()
        e.apply(clips, rolled, adv_flag=adv_flag, delta_clip=self.delta_clip)
        e.forward()
        e.loss(labels, improve_loss=self.improve_loss, targeted=self.targeted, use_logits=self.use_logits,
               margin=self.margin, global_batch=self.global_batch, stack=self.stack)
        e.backward()
        self.grad.copy_(torch.roll(self.grad, -int(shift), dims=0))       # d/d(delta) = roll^-1 of d/d(rolled)
        if self.frame_mask is not None:
            self.grad.mul_(self.frame_mask)
        if self.world > 1:
            fdist.allreduce_sum_(self.comm, self.pg)
        e.update(self.delta, self.grad, self.m, self.v, self.step_count, self.beta0, self.beta1, self.beta2,
                 self.beta3, lr=self.lr if lr is None else lr, delta_clip=self.delta_clip, stack=self.stack)
        return self.scalars

    # ---- CUDA graph of the whole step ----------------------------------------------------------
    def capture(self, clips, labels, adv_flag=1.0, lr=None):
        """Capture one step (all libfav launches + the NCCL all-reduce) into a CUDA graph bound to the
        given device tensors; `replay()` re-runs it after the caller refreshed `clips`/`labels` in place.
        Removes ~150 launch gaps per step."""
        # warm-up outside capture (lazy cudaFuncSetAttribute etc.); the two warm-up steps must not move the attack:
        # perturbation, Adam moments and step counter are put back afterwards (the capture itself executes nothing)
        saved = (self.delta.clone(), self.m.clone(), self.v.clone(), self.step_count.clone())
        for _ in range(2):
            self.step(clips, labels, adv_flag=adv_flag, lr=lr)
        torch.cuda.synchronize(self.device)
        for dst, src in zip((self.delta, self.m, self.v, self.step_count), saved):
            dst.copy_(src)
        torch.cuda.synchronize(self.device)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self.step(clips, labels, adv_flag=adv_flag, lr=lr)
        return self._graph

    def replay(self):
        self._graph.replay()
        return self.scalars

    # ---- end-to-end path: host buffers in, host scalars out ----------------------------------
    def _ensure_staging(self, like):
        if self._stage is None and like is None:
            raise RuntimeError("capture_staged: prefetch both staging slots first")
        if self._stage is None:
            self._stage = [torch.empty(like.shape, dtype=like.dtype, device=self.device) for _ in range(2)]
            self._stage_labels = [torch.empty((self.B,), dtype=torch.int64, device=self.device) for _ in range(2)]
            self._stage_events = [torch.cuda.Event() for _ in range(2)]
            self._done_events = [torch.cuda.Event() for _ in range(2)]
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._host_scalars = [torch.empty(L.S_COUNT, dtype=torch.float32).pin_memory() for _ in range(2)]
            self._slot = 0

    def prefetch(self, host_clips, host_labels):
        """Start the host->device copy of the NEXT batch (pinned host memory) on the copy stream."""
        self._ensure_staging(host_clips)
        slot = self._slot
        cs = self._copy_stream
        # the buffer may still be read by the step that used it two steps ago
        cs.wait_event(self._done_events[slot])
        with torch.cuda.stream(cs):
            self._stage[slot].copy_(host_clips, non_blocking=True)
            self._stage_labels[slot].copy_(host_labels, non_blocking=True)
            self._stage_events[slot].record(cs)
        self._slot ^= 1
        return slot

    def capture_staged(self, adv_flag=1.0, lr=None):
        """CUDA graphs of the step bound to the two staging slots, so that `step_staged` replays ~110 launches as one
        graph.  Call once after both slots have been prefetched (the capture itself runs no step; the slots' current
        contents are only read if the caller replays)."""
        self._ensure_staging(None)
        torch.cuda.synchronize(self.device)
        self._staged_graphs = []
        for slot in range(2):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.step(self._stage[slot], self._stage_labels[slot], adv_flag=adv_flag, lr=lr)
            self._staged_graphs.append(g)
        self._staged_cfg = (adv_flag, lr)

    def step_staged(self, slot, adv_flag=1.0, lr=None):
        """Run one step on a prefetched batch and start the device->host read of its scalars.
        Returns the pinned host tensor and the event that marks it valid."""
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self._stage_events[slot])
        if getattr(self, "_staged_graphs", None) and self._staged_cfg == (adv_flag, lr):
            self._staged_graphs[slot].replay()
        else:
            self.step(self._stage[slot], self._stage_labels[slot], adv_flag=adv_flag, lr=lr)
        self._host_scalars[slot].copy_(self.scalars, non_blocking=True)
        self._done_events[slot].record(cur)
        return self._host_scalars[slot], self._done_events[slot]

    # ---- evaluation helpers (forward only) ---------------------------------------------------
    def predict(self, clips, adv_flag=1.0, shift=0):
        """softmax [B,K] for clean (adv_flag=0) or perturbed clips — the reference's
        `k_i3d(inputs, adv_flag)` (utils/kinetics_i3d_utils.py:210-212); `shift`: evaluate with roll(delta, shift)
        (cyclic perturbation)."""
        delta = self._applied_delta()
        if shift:
            delta = torch.roll(delta, int(shift), dims=0).contiguous()
        self.eng.apply(clips, delta, adv_flag=adv_flag, delta_clip=self.delta_clip)
        logits = self.eng.forward()
        return torch.softmax(logits, dim=-1)

    def evaluator(self):
        """The forward-only evaluation handle of this attack (same network, batch and device), created on first use."""
        if self._eval is None:
            e = self.eng
            self._eval = EvalEngine(e.B, e.T, e.H, e.W, e.K, e.device.index or 0, arch=self.arch)
            self._eval.load_weights(self._weights)
        return self._eval

    def eval_batch(self, clips, labels, clips_adv=None, n_clips=None, targeted=False, target_class=None,
                   exclude_misclassify=True, shift=0, with_loss=False, want_probs=False):
        """One validation batch through the fused evaluation pass (row f3): clean and perturbed clips in one forward,
        `evaluator().counts` += [miss, valid] on the device.  `shift`: the network sees roll(delta, shift) (cyclic
        perturbation); with_loss: the adversarial-loss scalars of the perturbed rows go to `evaluator().scalars`."""
        ev = self.evaluator()
        delta = self._applied_delta()
        if shift:
            delta = torch.roll(delta, int(shift), dims=0)
        loss = None
        if with_loss:
            loss = dict(improve_loss=self.improve_loss, targeted=self.targeted, use_logits=self.use_logits,
                        margin=self.margin, stack=self.stack)
        return ev.eval_batch(clips, delta.contiguous(), labels, clips_adv=clips_adv, n_clips=n_clips,
                             delta_clip=self.delta_clip, targeted=targeted, target_class=target_class,
                             exclude_misclassify=exclude_misclassify, loss=loss, want_probs=want_probs)

    def eval_counts(self, reset=True):
        """(miss, valid) accumulated by eval_batch since the last reset — one device-to-host read per validation pass."""
        ev = self.evaluator()
        miss, valid = (int(x) for x in ev.counts.cpu())
        if reset:
            ev.reset_counts()
        return miss, valid

    def adversarial_video(self, clips, as_uint8=False):
        """`adversarial_inputs_rgb` (fp32, utils/kinetics_i3d_utils.py:139-142) or its uint8 view
        ((adv+1.0)*127.5).astype(uint8) (utils/stats_and_plot/stats_plots.py:57)."""
        if as_uint8:
            out = torch.empty(clips.shape, dtype=torch.uint8, device=self.device)
            self.eng.apply(clips, self._applied_delta(), delta_clip=self.delta_clip, adv_u8=out)
        else:
            out = torch.empty(clips.shape, dtype=torch.float32, device=self.device)
            self.eng.apply(clips, self._applied_delta(), delta_clip=self.delta_clip, adv_f32=out)
        return out

    def close(self):
        """Drops the captured graphs first (they reference the engine's buffers and, in sharded runs, the process
        group's communicator: the group can only be destroyed cleanly once no graph holds its collectives)."""
        self._graph = None
        self._staged_graphs = None
        if torch.cuda.is_available():
            torch.cuda.synchronize(self.device)
        self.eng.close()
        if self._eval is not None:
            self._eval.close()
            self._eval = None


class SparseAttack:
    """Per-pixel attack (FLICKERING_ATTACK = False): the reference's `kinetics_i3d_L12` graph
    (utils/kinetics_i3d_utils.py:308-521, eps [T,224,224,3] initialised to 1e-8, no +-0.4 clip, loss =
    adv + beta_1 * loss_L12, i3d_adversarial_main_universal.py:133) and the torch stack with
    attack_type "L12" (pert_size [3,T,112,112], loss = adv + lambda_ * L12, model.py:169-175,211-214).
    One `step()` = apply -> forward -> loss -> backward to every pixel -> L1,2 gradient + Adam."""

    def __init__(self, weights, batch, frames, attack_cfg=None, num_classes=400, device=0, lr=1e-3, arch="i3d",
                 delta_clip=None, init=None, process_group=None, sharded=True):
        self.eng = FlickerEngine(batch, frames, None, None, num_classes, device, arch=arch)
        self.eng.load_weights(weights)
        self.eng.pixels_enable()
        # sharded universal attack: clips split over ranks, the per-pixel gradient [T,H,W,3] is sum-all-reduced
        # (54 MB for I3D, 2.4 MB for the 112x112 nets); one rank: no collective, nothing changes
        self.pg = process_group
        self.world = 1
        if sharded and torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size(process_group)
        self.global_batch = batch * self.world
        self.arch = arch
        self.device = self.eng.device
        self.B, self.T, self.H, self.W = batch, frames, self.eng.H, self.eng.W
        torch_stack = self.eng.torch_stack
        self.stack = L.FAV_STACK_TORCH if torch_stack else L.FAV_STACK_TF
        cfg = dict(attack_cfg or {})
        self.improve_loss = bool(cfg.get("IMPROVE_ADV_LOSS", True))
        self.targeted = bool(cfg.get("TARGETED_ATTACK", False))
        self.use_logits = bool(cfg.get("USE_LOGITS", False))
        self.margin = float(cfg.get("PROB_MARGIN", 0.05))
        # TF: loss = adv + beta_0 * (beta_1 * L12), beta_0 = LAMBDA (universal.py:130-135); torch: lambda_ * L12 (model.py:173)
        self.reg_weight = (float(cfg.get("LAMBDA", 1.0)) if torch_stack
                           else float(cfg.get("LAMBDA", 1.0)) * float(cfg.get("BETA_1", 0.5)))
        self.delta_clip = (0.2 if torch_stack else 0.0) if delta_clip is None else delta_clip
        self.lr = lr
        shape = (frames, self.H, self.W, 3)
        if init is None:
            # TF: constant 1e-8 (kinetics_i3d_utils.py:333); torch: U(-1,1)*1e-6 (model.py:71)
            if torch_stack:
                self.delta = (torch.rand(shape, device=self.device) * 2 - 1) * 1e-6
            else:
                self.delta = torch.full(shape, 1e-8, dtype=torch.float32, device=self.device)
        else:
            self.delta = init.to(self.device, torch.float32).contiguous()
        if self.world > 1:
            # the torch stack draws its initial perturbation at random (model.py:71): every rank must start from rank 0's
            src = 0 if process_group is None else torch.distributed.get_global_rank(process_group, 0)
            torch.distributed.broadcast(self.delta, src=src, group=process_group)
        self.m = torch.zeros_like(self.delta)
        self.v = torch.zeros_like(self.delta)
        # packed exchange buffer [T*H*W*3 gradient | FAV_S_COUNT scalars]: ONE all-reduce per step, like FlickerAttack
        n = self.delta.numel()
        self.comm = torch.zeros(n + L.S_COUNT, dtype=torch.float32, device=self.device)
        self.grad = self.comm[:n].view(shape)
        self.scalars = self.comm[n:]
        self.eng.scalars = self.scalars
        self.step_count = torch.zeros(1, dtype=torch.int64, device=self.device)

    def check_replicas(self):
        if self.world > 1 and not fdist.replicas_equal(self.delta, self.pg):
            raise RuntimeError("the perturbation differs between ranks: the replicas have diverged")

    def state_dict(self):
        return {"delta": self.delta.cpu(), "m": self.m.cpu(), "v": self.v.cpu(), "step": int(self.step_count.item())}

    def load_state_dict(self, sd):
        self.delta.copy_(sd["delta"])
        self.m.copy_(sd["m"])
        self.v.copy_(sd["v"])
        self.step_count.fill_(int(sd["step"]))

    @property
    def perturbation(self):
        """[T,H,W,3] (I3D, the reference's eps_rgb) or [3,T,H,W] (torch stack, Perturbation.perturbation)"""
        return self.delta.permute(3, 0, 1, 2) if self.eng.torch_stack else self.delta

    def step(self, clips_u8, labels, adv_flag=1.0, lr=None):
        e = self.eng
        e.apply_pixels(clips_u8, self.delta, adv_flag=adv_flag, delta_clip=self.delta_clip)
        e.forward()
        e.loss(labels, improve_loss=self.improve_loss, targeted=self.targeted, use_logits=self.use_logits,
               margin=self.margin, global_batch=self.global_batch if self.world > 1 else 0, stack=self.stack)
        e.backward_pixels(self.grad)
        if self.world > 1:
            # margin loss = sum over samples, CE = mean over the GLOBAL batch (fav_loss divides by global_batch): in
            # both cases the global gradient is the plain sum of the rank gradients; the L1,2 term is added once,
            # after the exchange, by update_pixels.  Scalars 0..3 (adversarial loss, fooled count, probability sums)
            # ride in the same buffer; the remaining scalar slots are rewritten by update_pixels after the exchange.
            fdist.allreduce_sum_(self.comm, self.pg)
        e.update_pixels(self.delta, self.grad, self.m, self.v, self.step_count, self.reg_weight,
                        delta_clip=self.delta_clip, lr=self.lr if lr is None else lr, stack=self.stack)
        return self.scalars

    def predict(self, clips_u8, adv_flag=1.0):
        self.eng.apply_pixels(clips_u8, self.delta, adv_flag=adv_flag, delta_clip=self.delta_clip)
        return torch.softmax(self.eng.forward(), dim=-1)

    def close(self):
        self.eng.close()
