"""Host-side mirror of the reference's torch attack stack (utils_cv/action_recognition/model.py) on top of
the libfav engine: same class names, constructor arguments and result layouts, so the drivers
`r2plus1d_main_universal_attack.py` / `r2plus1d_main_statistics_single_video_attack.py` read the same.

  Perturbation            model.py:58-130   (size [3,T,1,1] flickering / [3,T,H,W] sparse)
  Losses                  model.py:132-250  (+ .label_prob after a call)
  Adversarial_metrics     model.py:253-330
  VideoLearnerAdversarial model.py:337-347 (ctor), fit :460-628 / train_an_epoch :630-788,
                          fit_many_videos / single-video loop :791-1203

What differs on purpose: the frozen network never computes weight gradients; the clean
prediction of a clip is computed once, not every step; data arrives as iterables of (uint8 clips
[B,T,H,W,3], labels) — the decord / DataLoader pipeline is SURVEY §8(f1).  The network, the apply, the loss
gradient, the backward-to-input and Adam all run in libfav kernels; torch tensors are device memory only.
"""
import os
import time
from collections import OrderedDict

import numpy as np
import torch

from . import _lib as L
from . import dist as fdist
from .attack import FlickerAttack, SparseAttack
from .engine import op_loss

DEFAULT_MEAN = (0.43216, 0.394666, 0.37645)     # utils_cv/action_recognition/dataset.py:28
DEFAULT_STD = (0.22803, 0.22145, 0.216989)      # :29


class Perturbation:
    """model.py:58-130.  The tensor lives on the device; `forward([x, adversarial])` takes the uint8 clip batch
    [B,T,H,W,3] the engine consumes and returns the normalised adversarial input [B,3,T,H,W] the reference's
    module would hand to the network (the engine fuses it into its stem, so the drivers below never call it)."""

    def __init__(self, size, requires_grad=True, device="cuda", max_value=None, min_value=None, max_norm=1.0,
                 cyclic_pert=False):
        self.size = tuple(size)
        self.device = device
        self.requires_grad = requires_grad
        self.perturbation = (torch.rand(self.size, device=device) * 2 - 1) * 0.000001       # model.py:71
        mean, std = np.array(DEFAULT_MEAN), np.array(DEFAULT_STD)
        self.max_value = np.min((1 - mean) / std) if max_value is None else max_value        # :72-73
        self.min_value = np.max((0.0 - mean) / std) if min_value is None else min_value      # :74-75
        self.max_norm = max_norm
        self.dynamic_max_norm = max_norm
        self.cyclic_pert = cyclic_pert
        self._engine = None

    def bind(self, engine):
        self._engine = engine
        return self

    # engine layout <-> reference layout
    def as_engine(self):
        """[3,T,1,1] -> [T,3]; [3,T,H,W] -> [T,H,W,3] (contiguous float32)"""
        p = self.perturbation
        if p.shape[2] == 1 and p.shape[3] == 1:
            return p.reshape(3, -1).t().contiguous()
        return p.permute(1, 2, 3, 0).contiguous()

    def from_engine(self, t):
        if t.dim() == 2:
            self.perturbation = t.t().reshape(3, -1, 1, 1).contiguous()
        else:
            self.perturbation = t.permute(3, 0, 1, 2).contiguous()

    def forward(self, *input):
        x, adversarial = input[0]
        if self._engine is None:
            raise L.FavError("Perturbation.forward needs an engine: call .bind(engine) (there is no CPU fallback)")
        e = self._engine
        adv = torch.empty((e.B, 3, e.T, e.H, e.W), dtype=torch.float32, device=e.device)
        d = self.as_engine()
        if d.dim() == 2:
            e.apply(x, d, adv_flag=1.0 if adversarial else 0.0, delta_clip=self.dynamic_max_norm, adv_f32=adv)
        else:
            e.apply_pixels(x, d, adv_flag=1.0 if adversarial else 0.0, delta_clip=self.dynamic_max_norm, adv_f32=adv)
        return adv

    __call__ = forward

    def clamp_perturbation(self):
        return self.perturbation.clamp(-1.0 * self.dynamic_max_norm, self.dynamic_max_norm)   # model.py:98-101

    def convert_adversarial_video_zero_one(self, adv_vid):
        x = adv_vid.detach().cpu().numpy()
        return (x.transpose([0, 2, 3, 4, 1]) + np.array(DEFAULT_MEAN) / np.array(DEFAULT_STD)) * np.array(DEFAULT_STD)

    def apply_perturbation(self, x):
        return self.convert_adversarial_video_zero_one(self.forward([x, True]))               # model.py:109-112

    def metric_calc(self):
        thickness = self.perturbation.abs().mean() * 100.0
        roughness = (torch.roll(self.perturbation, 1, dims=1) - self.perturbation).abs().mean() * 100.0
        return thickness, roughness

    def init_perturbation(self, perturbation=(), requires_grad=True, device="cuda"):
        if len(perturbation) == 0:
            self.perturbation = (torch.rand(self.size, device=self.device) * 2 - 1) * 0.000001
        else:
            self.perturbation = torch.from_numpy(np.asarray(perturbation, dtype=np.float32)).to(self.device)

    def get_perturbation(self):
        return self.clamp_perturbation(), self.perturbation


class Losses:
    """model.py:132-250.  `__call__(labels, model_logits, model_prob, perturbation)` -> [loss, adv_loss, reg_loss];
    the adversarial term runs in libfav's loss kernel (torch selection rules), which also yields dloss/dlogits."""

    def __init__(self, beta_1=0.5, lambda_=1.0, targeted=False, target_class=None, margin=0.05, improve_loss=False,
                 logits=False, attack_type="flickering"):
        self.beta_1 = beta_1
        self.lambda_ = lambda_
        self.targeted = targeted
        self.target_class = target_class
        self.margin = margin
        self.logits = logits
        self.improve_loss = improve_loss
        self.attack_type = attack_type
        self.regularization_loss = (self.flickering_regularization_loss if attack_type == "flickering"
                                    else self.L12_regularization_loss)
        self.label_prob = None
        self.dlogits = None

    def flickering_regularization_loss(self, perturbation):                                   # model.py:198-209
        norm_reg = torch.mean(perturbation ** 2) + 1e-12
        right, left = torch.roll(perturbation, 1, dims=1), torch.roll(perturbation, -1, dims=1)
        diff = torch.mean((perturbation - right) ** 2) + 1e-12
        lap = torch.mean((-2 * perturbation + right + left) ** 2) + 1e-12
        return self.beta_1 * norm_reg + (1 - self.beta_1) * (diff + lap)

    def L12_regularization_loss(self, perturbation):                                          # model.py:211-214
        return torch.sum(torch.sqrt(torch.mean(perturbation ** 2, [0, 2, 3]))) + 1e-12

    def adv_loss(self, labels, model_logits, model_prob=None):
        lab = labels if not self.targeted else torch.full_like(labels, int(self.target_class))
        probs, dlogits, scalars = op_loss(model_logits.contiguous(), lab.contiguous(), improve_loss=self.improve_loss,
                                          targeted=self.targeted, use_logits=self.logits, margin=self.margin,
                                          stack=L.FAV_STACK_TORCH)
        self.label_prob = probs.gather(1, lab.view(-1, 1))
        self.dlogits = dlogits
        return scalars[L.S_ADV_LOSS]

    def __call__(self, labels, model_logits, model_prob, perturbation):
        reg_loss = self.regularization_loss(perturbation)
        adv_loss = self.adv_loss(labels, model_logits, model_prob)
        return [adv_loss + self.lambda_ * reg_loss, adv_loss, reg_loss]


class Adversarial_metrics:
    """model.py:253-330 (host logic on [B,K] tensors)."""

    def __init__(self, targeted=False, target_class=None):
        self.targeted = targeted
        self.target_class = target_class

    def accuracy_for_eval(self, output, ground_truth, topk=(1,), clean_pred=None):
        with torch.no_grad():
            batch_size = ground_truth.size(0)
            maxk = max(topk)
            _, pred = output.topk(maxk, 1, True, True)
            pred = pred.t()
            if self.targeted:
                correct = pred.eq(self.target_class).view(-1).float().sum(0, keepdim=True)
                return [correct[0].mul_(100.0 / batch_size)]
            correct = pred.eq(ground_truth.view(1, -1).expand_as(pred))
            _, pred_no_adv = clean_pred.topk(maxk, 1, True, True)
            correct_no_adv = pred_no_adv.t().eq(ground_truth.view(1, -1).expand_as(pred))
            miss = num = None
            for k in topk:
                miss = (torch.logical_not(correct[:k]) * correct_no_adv[:k]).view(-1).float().sum(0, keepdim=True)
                num = correct_no_adv[:k].view(-1).float().sum()
            return miss, num

    def adversarial_metric(self, perturbation):
        thickness = perturbation.abs().mean() * 100.0
        roughness = (torch.roll(perturbation, 1, dims=1) - perturbation).abs().mean() * 100.0
        return thickness, roughness


def step_lr_schedule(lr, epochs, start_epoch=1, lr_gamma=0.1, lr_step_size=None):
    """{epoch: learning rate} of the reference's `torch.optim.lr_scheduler.StepLR(optimizer, step_size, gamma)` stepped
    once after every epoch of `range(start_epoch, epochs + 1)` (model.py:495-497, 571-573, 587-608).  The schedule is
    produced by torch's own scheduler on a dummy optimizer, so its rounding and step rules are the reference's."""
    if lr_step_size is None:
        lr_step_size = np.ceil(2 / 3 * epochs)
    dummy = torch.optim.Adam([torch.zeros(1, requires_grad=True)], lr=lr)
    sched = torch.optim.lr_scheduler.StepLR(dummy, step_size=int(lr_step_size), gamma=lr_gamma)
    out = {}
    for e in range(start_epoch, epochs + 1):
        out[e] = float(dummy.param_groups[0]["lr"])
        dummy.step()
        sched.step()
    return out


class VideoLearnerAdversarial:
    """model.py:337-347.  `weights` is the torchvision state_dict of `base_model` (the reference downloads the
    pretrained one, :421; there is no network here)."""

    def __init__(self, dataset=None, num_classes=400, base_model="r2plus1d_18", sample_length=16, cyclic_pert=False,
                 l_inf_pert_norm=1.0, attack_type="flickering", labaels_id_to_text=None, weights=None, batch_size=8,
                 device=0, seed=None):
        if base_model not in ("r3d_18", "mc3_18", "r2plus1d_18"):
            raise ValueError(f"base_model {base_model!r}: the engine implements r3d_18 / mc3_18 / r2plus1d_18")
        if weights is None:
            raise ValueError("weights (a torchvision state_dict) are required: there is no checkpoint download")
        self.results = []
        self.num_classes = num_classes
        self.attack_type = attack_type
        self.label_id_to_text = labaels_id_to_text
        self.dataset = dataset
        self.sample_length = sample_length
        self.model_name = base_model
        self.batch_size = batch_size
        self._weights, self._device = weights, device
        self._rng = np.random.RandomState(seed)          # random shifts of the cyclic perturbation attack
        pert_size = (3, sample_length, 1, 1) if attack_type == "flickering" else (3, sample_length, 112, 112)
        self.pert_model = Perturbation(size=pert_size, device=torch.device("cuda", device), max_norm=l_inf_pert_norm,
                                       cyclic_pert=cyclic_pert)
        self._atk = None

    def _attack(self, lr, loss_params_dict, batch, sharded=True):
        cfg = {"LAMBDA": loss_params_dict["lambda_"], "BETA_1": loss_params_dict["beta_1"],
               "TARGETED_ATTACK": loss_params_dict["targeted_attack"], "IMPROVE_ADV_LOSS": loss_params_dict["improve_loss"],
               "USE_LOGITS": loss_params_dict["use_logits"], "PROB_MARGIN": 0.05}
        mean, std = np.array(DEFAULT_MEAN), np.array(DEFAULT_STD)
        if not (np.isclose(self.pert_model.max_value, np.min((1 - mean) / std)) and
                np.isclose(self.pert_model.min_value, np.max((0.0 - mean) / std))):
            # the apply kernel clamps to the reference's default scalar bounds (model.py:72-75); other bounds are refused
            # rather than silently ignored (the attack mains never pass them)
            raise NotImplementedError("Perturbation max_value / min_value other than the defaults")
        if self.pert_model.cyclic_pert and self.attack_type != "flickering":
            raise NotImplementedError("cyclic_pert is built for the flickering attack only")
        if self.attack_type == "flickering":
            atk = FlickerAttack(self._weights, batch, self.sample_length, cfg, num_classes=self.num_classes,
                                device=self._device, lr=lr, arch=self.model_name,
                                delta_clip=self.pert_model.dynamic_max_norm, sharded=sharded)
            atk.delta.copy_(self.pert_model.as_engine())
            if atk.world > 1:      # Perturbation draws U(-1,1)*1e-6 per process (model.py:71): all ranks take rank 0's
                torch.distributed.broadcast(atk.delta, src=0)
        else:
            atk = SparseAttack(self._weights, batch, self.sample_length, cfg, num_classes=self.num_classes,
                               device=self._device, lr=lr, arch=self.model_name,
                               delta_clip=self.pert_model.dynamic_max_norm, init=self.pert_model.as_engine(),
                               sharded=sharded)
        self.pert_model.bind(atk.eng)
        return atk

    # ---- resume (r2plus1d_main_universal_attack.py:197-216) -------------------------------------------------
    def resume_from(self, model_dir, model_name=None, init_pert=True, continue_train=True):
        """The two restart switches of the reference's main.  INIT_PERT_FROM_LAST_CKPT (`init_pert`): the perturbation is
        re-read from `valid/perturbation` of the newest `{model_name}_{epoch:03d}.npy` in `model_dir`; CONTINUE_TRAIN
        (`continue_train`): the returned `start_epoch` continues that file's epoch number (1 when there is none).  When the
        epoch also left a `.state.npz` next to it (written by `fit(save_model=True)`), the following `fit(start_epoch=...)`
        restores the Adam moments and step as well, so that the continued run is bit-identical to an uninterrupted one —
        the reference restarts its optimizer here."""
        import glob
        files = sorted(glob.glob(os.path.join(model_dir, "{}_[0-9][0-9][0-9].npy".format(model_name or self.model_name))))
        if not files:
            return 1
        last = files[-1]
        if init_pert:
            pert = np.load(last, allow_pickle=True)[-1]["valid/perturbation"]
            self.pert_model.init_perturbation(pert)
        self._resume_state = last[:-4] + ".state.npz"
        return int(last[:-4].split("_")[-1]) + 1 if continue_train else 1

    def _cyclic_shift(self):
        """Perturbation.forward draws `np.random.randint(0, T)` per adversarial forward when cyclic_pert (model.py:91-92)"""
        if not self.pert_model.cyclic_pert:
            return 0
        rng = getattr(self, "_rng", None) or np.random
        return int(rng.randint(0, self.sample_length))

    def _sync_pert(self, atk):
        self.pert_model.from_engine(atk.delta)

    # ---- universal attack: fit (model.py:460-628) with train_an_epoch (:630-788) --------------------------
    def fit(self, lr, epochs, model_dir="checkpoints", model_name=None, momentum=0.95, weight_decay=0.0001,
            mixed_prec=False, use_one_cycle_policy=False, warmup_pct=0.3, lr_gamma=0.1, lr_step_size=None, grad_steps=2,
            save_model=False, loss_params_dict=None, devices_ids=None, start_epoch=1, *, train_batches=None,
            valid_batches=None):
        """Same signature as the reference's `fit` (model.py:460-478).  Epochs run from `start_epoch` to `epochs`
        INCLUSIVE (:587) and epoch e is saved as `{model_name}_{e:03d}.npy` (pickled list of per-epoch OrderedDicts,
        :619-623); the learning rate follows the reference's StepLR (step `lr_step_size`, default ceil(2/3 * epochs),
        factor `lr_gamma`, stepped once per epoch and restarted at `lr` on every call, :495-497, 571-573, 608).
        `momentum`, `weight_decay` and `grad_steps` are unused by the reference's Adam loop as well; `mixed_prec` and
        `devices_ids` have no meaning here (bf16 tensor-core engine, one process per GPU).
        train_batches / valid_batches (keyword only): callables returning an iterable of (uint8 clips [B,T,H,W,3]
        DEVICE, labels [B] DEVICE) per epoch; by default the `dataset` given to the constructor supplies them
        (`video_dataset.VideoDataset.train_batches` / `.test_batches`, the reference's `dataset.train_dl` / `.test_dl`,
        model.py:506-507).  Returns `self.results`."""
        if use_one_cycle_policy:
            raise NotImplementedError("use_one_cycle_policy: only the StepLR schedule of the attack drivers is built")
        from_dataset = train_batches is None or valid_batches is None
        if from_dataset:
            if self.dataset is None:
                raise ValueError("fit needs train_batches / valid_batches or a dataset")
            if self.dataset.batch_size != self.batch_size or self.dataset.sample_length != self.sample_length:
                raise ValueError("dataset batch_size / sample_length differ from the learner's")
            train_batches = train_batches or self.dataset.train_batches
            valid_batches = valid_batches or self.dataset.test_batches
        lr_schedule = step_lr_schedule(lr, epochs, start_epoch, lr_gamma, lr_step_size)
        lp = dict(loss_params_dict)
        metric = Adversarial_metrics(lp["targeted_attack"], lp.get("target_class_id"))
        atk = self._attack(lr, lp, self.batch_size)
        self._atk = atk
        if atk.world > 1 and from_dataset and hasattr(self.dataset, "set_shard"):
            # the gradient is SUMMED over ranks: every rank must see different clips (else the data term is world x too
            # large against the regulariser) and the same number of them (one collective per step)
            self.dataset.set_shard(torch.distributed.get_rank(atk.pg), atk.world, atk.device, atk.pg)
        os.makedirs(model_dir, exist_ok=True)
        model_name = model_name or self.model_name
        target = lp.get("target_class_id")
        state_path = getattr(self, "_resume_state", None)
        if state_path and start_epoch > 1 and os.path.exists(state_path):
            from . import checkpoint as fckpt       # continue exactly: perturbation, Adam moments and Adam step
            fckpt.restore_checkpoint(state_path, atk)
            self._sync_pert(atk)
        self._resume_state = None
        for e in range(start_epoch, epochs + 1):
            lr_e = lr_schedule[e]
            result = OrderedDict()
            for phase, batches in (("train", train_batches), ("valid", valid_batches)):
                t0 = time.time()
                miss_rate = total = 0.0
                loss_sum = n_seen = 0.0
                # caller-supplied sharded sources may differ in length per rank: the training pass runs in lockstep
                it = fdist.lockstep(batches(), atk.world, atk.device, atk.pg) if phase == "train" else batches()
                # validation (model.py:697-713 runs model([x, False]) and model([x, True]) per batch): the fused
                # evaluation pass forwards both versions at once and keeps the miss / valid counters on the device
                fused_valid = phase == "valid" and not lp["targeted_attack"]
                if fused_valid:
                    atk.eval_counts(reset=True)
                    reg_valid = None
                for clips, labels in it:
                    lab = labels if not lp["targeted_attack"] else torch.full_like(labels, int(target))
                    if fused_valid:
                        atk.eval_batch(clips, labels, shift=self._cyclic_shift(), with_loss=True)
                        if reg_valid is None:          # the perturbation does not change during validation
                            self._sync_pert(atk)
                            reg_valid = float(Losses(lp["beta_1"], lp["lambda_"], attack_type=self.attack_type)
                                              .regularization_loss(self.pert_model.get_perturbation()[0]))
                        loss = float(atk.evaluator().scalars[L.S_ADV_LOSS]) + lp["lambda_"] * reg_valid
                        loss_sum += loss * labels.numel()
                        n_seen += labels.numel()
                        continue
                    clean = atk.predict(clips, adv_flag=0.0).clone()
                    shift = self._cyclic_shift()
                    if phase == "train":
                        sc = atk.step_rolled(clips, lab, shift, lr=lr_e) if shift else atk.step(clips, lab, lr=lr_e)
                        loss = float(sc[L.S_TOTAL_LOSS])
                        # scores of the delta the step was computed with are still in the engine
                        adv_logits = atk.eng.logits.clone()
                    else:
                        atk.predict(clips, adv_flag=1.0, shift=shift) if shift else atk.predict(clips, adv_flag=1.0)
                        sc = atk.eng.loss(lab, improve_loss=lp["improve_loss"], targeted=lp["targeted_attack"],
                                          use_logits=lp["use_logits"], margin=0.05, stack=L.FAV_STACK_TORCH)
                        self._sync_pert(atk)
                        reg = Losses(lp["beta_1"], lp["lambda_"], attack_type=self.attack_type).regularization_loss(
                            self.pert_model.get_perturbation()[0])
                        loss = float(sc[L.S_ADV_LOSS]) + lp["lambda_"] * float(reg)
                        adv_logits = atk.eng.logits.clone()
                    out = metric.accuracy_for_eval(adv_logits, labels, topk=(1,), clean_pred=clean)
                    if lp["targeted_attack"]:
                        miss_rate += float(out[0]) * labels.numel() / 100.0
                        total += labels.numel()
                    else:
                        miss_rate += float(out[0][0])
                        total += float(out[1])
                    loss_sum += loss * labels.numel()
                    n_seen += labels.numel()
                if fused_valid:
                    miss_rate, total = (float(x) for x in atk.eval_counts(reset=True))
                self._sync_pert(atk)
                atk.check_replicas()          # sharded run: the replicated perturbation must be identical on all ranks
                pert = self.pert_model.get_perturbation()[0].detach().cpu().numpy()
                # sharded epochs (one process per GPU, each over its shard of the clips): metrics over all ranks
                if atk.world > 1:
                    miss_rate, total, loss_sum, n_seen = fdist.sum_counts((miss_rate, total, loss_sum, n_seen),
                                                                         device=atk.device, group=atk.pg)
                result[f"{phase}/time"] = time.time() - t0
                result[f"{phase}/loss"] = loss_sum / max(n_seen, 1.0)
                result[f"{phase}/fooling_ratio"] = miss_rate / max(total, 1.0)
                result[f"{phase}/pert_thickness"] = np.abs(pert).mean()
                result[f"{phase}/pert_roughness"] = np.abs(np.roll(pert, 1, 1) - pert).mean()
                result[f"{phase}/inf_norm"] = np.abs(pert).max()
                result[f"{phase}/perturbation"] = pert
            self.results.append(result)
            if save_model and fdist.is_writer(atk.world):
                np.save(os.path.join(model_dir, "{}_{:03d}.npy".format(model_name, e)),
                        np.array(self.results, dtype=object), allow_pickle=True)
                sd = atk.state_dict()          # sidecar for an exact restart (resume_from)
                np.savez(os.path.join(model_dir, "{}_{:03d}.state.npz".format(model_name, e)), delta=sd["delta"].numpy(),
                         m=sd["m"].numpy(), v=sd["v"].numpy(), adam_step=np.int64(sd["step"]), step=np.int64(e))
        return self.results

    # ---- single-video attack (model.py:918-1203) --------------------------------------------------------------
    def fit_single_video(self, lr, n_iter, clip_u8, label, video_name="video", class_name=None, model_dir=None,
                         loss_params_dict=None, max_restarts=4, restart_after=3000, reuse_attack=None,
                         reset_optimizer=False):
        """One clip until `step >= n_iter and adversarial`; every `restart_after` steps without success
        dynamic_max_norm *= 1.3 (at most `max_restarts` times, model.py:1061-1066).  Returns the result dict the
        reference saves as `{vid}_@{class}.npy` (:1194-1203), or None when the clean clip is misclassified.
        `reuse_attack`: an attack object of an earlier call (same loss parameters) to continue with — engine, packed
        weights AND Adam state are kept, only the perturbation is re-read from `pert_model`; this is what the
        reference's `fit_many_videos` does with its single optimizer (model.py:868, 949-952)."""
        lp = dict(loss_params_dict)
        if reuse_attack is None:
            atk = self._attack(lr, lp, 1, sharded=False)      # single-video attacks are per-rank replicas (no collective)
        else:
            atk = reuse_attack
            atk.lr, atk.delta_clip = lr, self.pert_model.dynamic_max_norm
            atk.delta.copy_(self.pert_model.as_engine())
            if reset_optimizer:               # a fresh Adam per video instead of the reference's carried-over moments
                atk.m.zero_()
                atk.v.zero_()
                atk.step_count.zero_()
            self.pert_model.bind(atk.eng)
        self._atk = atk
        clips = clip_u8.reshape(1, *clip_u8.shape[-4:]).contiguous()
        labels = torch.as_tensor([int(label)], dtype=torch.int64, device=atk.device)
        clean = atk.predict(clips, adv_flag=0.0).clone()
        if int(clean.argmax()) != int(label):
            return None
        tgt = labels if not lp["targeted_attack"] else torch.full_like(labels, int(lp["target_class_id"]))
        res = {"loss/total": [], "loss/adv_loss": [], "loss/reg_loss": [], "perturbation/thickness": [],
               "perturbation/roughness": [], "perturbation/inf_norm": 0.0, "perturbation": [],
               "prob_clean_input": atk.eng.logits.clone().cpu().numpy(),      # `outputs_no_adv`: the clean LOGITS (:1193)
               "label": np.asarray([int(label)]), "is_adversarial": []}
        step = new_chance = 0
        is_adv = False
        while step < n_iter or not is_adv:
            if step > restart_after:
                new_chance += 1
                self.pert_model.dynamic_max_norm *= 1.3
                atk.delta_clip = self.pert_model.dynamic_max_norm
                step = 0
            if new_chance == max_restarts:
                break
            shift = self._cyclic_shift()
            sc = (atk.step_rolled(clips, tgt, shift) if shift else atk.step(clips, tgt)).clone()
            pred = int(atk.eng.logits.argmax())
            is_adv = (pred == int(lp["target_class_id"])) if lp["targeted_attack"] else (pred != int(label))
            self._sync_pert(atk)
            pert = self.pert_model.get_perturbation()[0].detach().cpu().numpy()
            sc = sc.cpu()
            res["loss/total"].append(float(sc[L.S_TOTAL_LOSS]))
            res["loss/adv_loss"].append(float(sc[L.S_ADV_LOSS]))
            res["loss/reg_loss"].append(float(sc[L.S_TOTAL_LOSS] - sc[L.S_ADV_LOSS]))
            res["perturbation/thickness"].append(np.abs(pert).mean())
            res["perturbation/roughness"].append(np.abs(np.roll(pert, 1, 1) - pert).mean())
            res["perturbation/inf_norm"] = np.abs(pert).max()          # the reference keeps the final value only (:1199)
            res["perturbation"].append(pert)
            res["is_adversarial"].append(bool(is_adv))
            step += 1
        if model_dir is not None:
            os.makedirs(model_dir, exist_ok=True)
            np.save(os.path.join(model_dir, "{}_@{}.npy".format(video_name, class_name if class_name else label)), res,
                    allow_pickle=True)
        return res

    # ---- single-video attack over a dataset (model.py:789-979) -----------------------------------------------
    def fit_many_videos(self, lr, epochs=1, model_dir="checkpoints", model_name=None, momentum=0.95,
                        weight_decay=0.0001, mixed_prec=False, use_one_cycle_policy=False, warmup_pct=0.3, lr_gamma=0.1,
                        lr_step_size=None, grad_steps=2, save_model=False, loss_params_dict=None, devices_ids=None, *,
                        n_iter=3000, videos=None, max_restarts=4, restart_after=3000, share_optimizer_state=True):
        """`fit_many_videos` (same positional signature as the reference, model.py:789-806; the optimiser / schedule
        arguments are accepted and unused — the reference creates a scheduler here but never steps it in the
        single-video loop): one single-video attack per video of the dataset's training split.  For each video the
        result goes to `{model_dir}/{video}_@{class_name}.npy` (spaces in the class name replaced by `_`); a video whose
        file already holds a successful attack is skipped, one whose file holds `None` (claimed by another run, or
        clean-misclassified) too; with `save_model` a `None` placeholder is written before the attack starts
        (model.py:925-946).  Before every video the perturbation is re-drawn as U(-1,1) * 0.005 and
        `dynamic_max_norm` reset (:949-952).  `videos`: optional iterable of (uint8 DEVICE clip [T,H,W,3], label, path)
        replacing the dataset.  With torch.distributed initialised the videos are dealt round-robin to the ranks
        (replicas only: no collective).  `share_optimizer_state` (default True, the reference's behaviour): the
        Adam moments and step count carry over from one video to the next because the reference re-draws the parameter
        under ONE optimizer (model.py:868, 946-948); False starts every video with a fresh Adam.  Returns
        {video name: result dict or None}."""
        lp = dict(loss_params_dict)
        os.makedirs(model_dir, exist_ok=True)
        if videos is None:
            if self.dataset is None:
                raise ValueError("fit_many_videos needs a dataset or `videos`")
            videos = (self.dataset[i] for i in self.dataset.train_range)
        rank, world = 0, 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            rank, world = torch.distributed.get_rank(), torch.distributed.get_world_size()
        out, shared = {}, None
        for vid_num, (clip, target, vid_path) in enumerate(videos):
            if vid_num % world != rank:
                continue
            target = int(target)
            vid_name = str(vid_path).split("/")[-1]
            class_name = str(self.label_id_to_text[target]) if self.label_id_to_text is not None else str(target)
            dest_path = os.path.join(model_dir, "{}_@{}.npy".format(vid_name, class_name.replace(" ", "_")))
            if os.path.exists(dest_path):
                prev = np.load(dest_path, allow_pickle=True).tolist()
                if prev is None or np.array(prev["is_adversarial"]).any():
                    continue
            elif save_model:
                np.save(dest_path, None)
            self.pert_model.perturbation = (torch.rand(self.pert_model.size, device=self.pert_model.device) * 2 - 1) * 0.005
            self.pert_model.dynamic_max_norm = self.pert_model.max_norm
            res = self.fit_single_video(lr, n_iter, clip, target, loss_params_dict=lp, max_restarts=max_restarts,
                                        restart_after=restart_after, reuse_attack=shared,
                                        reset_optimizer=not share_optimizer_state)
            shared = self._atk             # one engine and ONE optimizer state for all videos, like the reference
            out[vid_name] = res
            if res is not None and save_model:
                np.save(dest_path, res, allow_pickle=True)
        return out
