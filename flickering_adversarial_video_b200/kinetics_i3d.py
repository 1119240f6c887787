"""Host-side mirror of the reference's TF attack operator surface (utils/kinetics_i3d_utils.py:76-307),
eager instead of graph-mode: the same constructor keywords, attribute names and methods, backed by
the libfav engine.

Reference execution model: drivers build a loss from the object's tensor handles, create Adam on
`eps_rgb`, then `sess.run(fetches, feed_dict)` (i3d_adversarial_main_single_video_npy.py:37-84,
211-215).  Here `improve_adversarial_loss(...)` / `ce_adversarial_loss(...)` select the loss,
`train_step(...)` is one `sess.run([train_op, loss, ...])`, and the attribute names the drivers fetch
(`softmax`, `model_logits`, `norm_reg`, `thickness`, `to_min_prob`, …) are properties holding the
values of the last step as numpy arrays, exactly what `sess.run` returned.
"""
import os

import numpy as np
import torch

from . import _lib as L
from . import dist as fdist
from .attack import FlickerAttack

_IMAGE_SIZE = 224
_SAMPLE_VIDEO_FRAMES = 90     # utils/kinetics_i3d_utils.py:12 (a parameter here: `frames=`)
NUM_CLASSES = 400
_LABEL_MAP_PATH = "data/label_map.txt"


def load_kinetics_classes(eval_type="rgb", label_map_path=_LABEL_MAP_PATH):
    """utils/kinetics_i3d_utils.py:67-74 (rgb600 was never functional in the reference: :70)."""
    if eval_type == "rgb600":
        raise ValueError("rgb600 is not supported (the reference's _LABEL_MAP_PATH_600 is undefined)")
    if not os.path.exists(label_map_path):
        return [f"class_{i:03d}" for i in range(NUM_CLASSES)]
    return [x.strip() for x in open(label_map_path)]


def load_weights(ckpt_path):
    """Weights as {tf variable name: float32 array}: a TF1 checkpoint prefix (`<prefix>.index` +
    `<prefix>.data-*`, what the reference's Saver restores — utils/kinetics_i3d_utils.py:41-62; read by ckpt.py) or
    an .npz with the same variable names (`RGB/inception_i3d/...`)."""
    if isinstance(ckpt_path, dict):
        return ckpt_path
    if ckpt_path and os.path.exists(ckpt_path + ".index"):
        from .ckpt import read_tf_checkpoint
        return {k: v.astype(np.float32) for k, v in read_tf_checkpoint(ckpt_path).items()
                if k.startswith("RGB/inception_i3d/")}
    if ckpt_path and os.path.exists(ckpt_path) and ckpt_path.endswith(".npz"):
        with np.load(ckpt_path) as z:
            return {k: z[k].astype(np.float32) for k in z.files}
    raise FileNotFoundError(
        f"checkpoint '{ckpt_path}' not found (neither a TF checkpoint prefix nor an .npz of TF variable names); "
        "pass weights=<dict> (e.g. synthetic.i3d_weights()) — there is no network to fetch the Kinetics ckpt")


def float_clip_to_u8(t, strict=True):
    """Exact inverse of the reference's clip normalisation x = u8/128 - 1 (utils/pre_process_rgb_flow.py:234):
    u8 = (x + 1) * 128, which is exact in fp32 for every grid value.  Returns the uint8 tensor, or — when some value
    is off the grid or outside [0, 255] — raises ValueError (strict) / returns None."""
    u = (t + 1.0) * 128.0
    r = u.round()
    ok = bool(torch.equal(u, r)) and float(r.min()) >= 0.0 and float(r.max()) <= 255.0
    if not ok:
        if strict:
            raise ValueError("float clip is not on the reference's grid uint8/128 - 1 (utils/pre_process_rgb_flow.py:234); "
                             "pass the uint8 video")
        return None
    return r.to(torch.uint8)


class _KineticsGraph:
    """What the two reference graph classes share (utils/kinetics_i3d_utils.py:76-307 and :308-521 repeat it verbatim):
    the attribute handles of the last run, the loss selection, the clean / perturbed forward."""

    def _init_common(self, ckpt_path, batch_size, frames, default_adv_flag_c, cyclic_flag_default_c, label_map_path,
                     weights, seed, rgb_input, labels):
        self.ckpt_path = ckpt_path
        self.batch_size = batch_size
        self.frames = frames
        self.adv_flag = float(default_adv_flag_c)
        self.cyclic_flag = float(cyclic_flag_default_c)
        self.kinetics_classes = load_kinetics_classes(label_map_path=label_map_path)
        self._rng = np.random.RandomState(seed)
        self._loss_cfg = dict(improve=True, margin=0.05, targeted=False, logits=False)
        self._last = {}
        self.rgb_input = rgb_input
        self.labels = labels
        return load_weights(weights if weights is not None else ckpt_path)

    # ---- the handles the drivers read (values of the last run) ------------------------------
    @property
    def eps_rgb(self):
        return self._atk.perturbation.detach().cpu().numpy()

    perturbation = eps_rgb

    def __getattr__(self, name):
        last = self.__dict__.get("_last", {})
        if name in last:
            return last[name]
        raise AttributeError(name)

    def get_kinetics_classes(self):
        return self.kinetics_classes

    # ---- loss selection (utils/kinetics_i3d_utils.py:253-307 / :467-521) ---------------------
    def improve_adversarial_loss(self, margin=0.05, targeted=False, logits=False):
        self._loss_cfg = dict(improve=True, margin=float(margin), targeted=bool(targeted), logits=bool(logits))
        return "adversarial_loss_total"

    def ce_adversarial_loss(self, targeted=False):
        self._loss_cfg = dict(improve=False, margin=0.05, targeted=bool(targeted), logits=False)
        return "adversarial_loss_total"

    def _configure(self):
        a, c = self._atk, self._loss_cfg
        a.improve_loss, a.margin, a.targeted, a.use_logits = c["improve"], c["margin"], c["targeted"], c["logits"]

    def _prob_handles(self, probs, lab_np):
        """label_prob, max_non_label_prob, to_min_prob, to_max_prob (utils/kinetics_i3d_utils.py:152-167)"""
        label_prob = probs[np.arange(probs.shape[0]), lab_np]
        onehot = np.eye(probs.shape[1], dtype=np.float32)[lab_np]
        max_non_label_prob = (probs - onehot).max(-1)
        tmin, tmax = (max_non_label_prob, label_prob) if self._loss_cfg["targeted"] else (label_prob, max_non_label_prob)
        return label_prob, max_non_label_prob, tmin, tmax

    def __call__(self, inputs, adv_flag=0):
        """softmax for `inputs` (utils/kinetics_i3d_utils.py:210-212 / :424-426)."""
        clips = self._to_device_clip(inputs)
        self._last_clips = clips
        self.prob = self._atk.predict(clips, adv_flag=float(adv_flag)).cpu().numpy()
        return self.prob

    def reset(self):
        """sess.run(eps_rgb.initializer); sess.run(tf.variables_initializer(optimizer.variables()))"""
        self._atk.reset()

    def close(self):
        self._atk.close()


class kinetics_i3d(_KineticsGraph):
    """Drop-in for ki3du.kinetics_i3d(ckpt_path, batch_size, init_model, rgb_input, labels,
    cyclic_flag_default_c, cyclic_pert_flag_default_c, default_adv_flag_c)."""

    flickering = True

    def __init__(self, ckpt_path="data/checkpoints/rgb_imagenet/model.ckpt", batch_size=1, init_model=True,
                 rgb_input=None, labels=None, cyclic_flag_default_c=0.0, cyclic_pert_flag_default_c=0.0,
                 default_adv_flag_c=1.0, frames=_SAMPLE_VIDEO_FRAMES, weights=None, device=0,
                 label_map_path=_LABEL_MAP_PATH, seed=0, sharded=True, frame_range=None):
        w = self._init_common(ckpt_path, batch_size, frames, default_adv_flag_c, cyclic_flag_default_c, label_map_path,
                              weights, seed, rgb_input, labels)
        self.cyclic_pert_flag = float(cyclic_pert_flag_default_c)
        # sharded=False: a per-rank replica (single-video attacks under torchrun), never joins a collective
        # frame_range = (_IND_START, _IND_END) of utils/kinetics_i3d_utils.py:14-15 (module constants there); None = all frames
        self._atk = FlickerAttack(w, batch_size, frames, {}, device=device, sharded=sharded, frame_range=frame_range)
        self.device = self._atk.device

    def _clips_of_last_run(self):
        clips = self.__dict__.get("_last_clips")
        if clips is None:
            raise AttributeError("no input has been run yet (the reference fetches this handle with a feed_dict)")
        return clips

    @property
    def softmax_clean(self):
        """softmax of the unperturbed input of the last run (utils/kinetics_i3d_utils.py:146-149), evaluated on demand"""
        return self._atk.predict(self._clips_of_last_run(), adv_flag=0.0).cpu().numpy()

    @property
    def adversarial_inputs_rgb(self):
        """clip(input + adv_flag * perturbation, -1, 1) for the input of the last run (:139-142), fp32 [B,T,224,224,3]"""
        return self._atk.adversarial_video(self._clips_of_last_run()).cpu().numpy()

    def set_perturbation(self, value):
        self._atk.delta.copy_(torch.as_tensor(np.asarray(value), dtype=torch.float32).reshape(self.frames, 3))

    def _to_device_clip(self, inputs):
        """uint8 clips go to the device as they are.  Float clips that lie on the reference's grid u/128 - 1
        (parse_example_uint8, utils/pre_process_rgb_flow.py:234 — every clip its loaders produce) are turned back
        into that uint8 video exactly (4x less upload, the coalesced uint8 apply kernel); any other float clip takes
        the engine's fp32 input path unchanged."""
        t = torch.as_tensor(inputs)
        if t.is_cuda and t.dtype in (torch.uint8, torch.float32):
            return t.reshape(self.batch_size, self.frames, _IMAGE_SIZE, _IMAGE_SIZE, 3).contiguous()
        if t.dtype != torch.uint8:
            t = t.to(torch.float32)
            u8 = float_clip_to_u8(t, strict=False)
            t = t if u8 is None else u8
        t = t.reshape(self.batch_size, self.frames, _IMAGE_SIZE, _IMAGE_SIZE, 3)
        return t.to(self.device, non_blocking=True).contiguous()

    # ---- sess.run equivalents ----------------------------------------------------------------
    def train_step(self, inputs, labels, learning_rate=1e-3, beta_0=1.0, beta_1=0.1, beta_2=0.1, beta_3=0.1,
                   cyclic_flag=None, cyclic_pert_flag=None, adv_flag=None):
        """One `sess.run([train_op, loss, adversarial_loss, regularizer_loss, norm_reg, diff_norm_reg,
        laplacian_norm_reg, thickness, roughness, prob_to_max, prob_to_min], feed_dict)`.
        The fetched values describe the perturbation BEFORE the update, the returned softmax is
        evaluated AFTER it (the reference's second sess.run, single_video_npy.py:217) only on request
        via `softmax_after_update`."""
        self._configure()
        a = self._atk
        clips = self._to_device_clip(inputs)
        lab = torch.as_tensor(np.asarray(labels), dtype=torch.int64).reshape(-1).to(self.device)
        cyc = self.cyclic_flag if cyclic_flag is None else float(cyclic_flag)
        cycp = self.cyclic_pert_flag if cyclic_pert_flag is None else float(cyclic_pert_flag)
        flag = self.adv_flag if adv_flag is None else float(adv_flag)
        shift_p = 0
        if cyc:       # tf.roll(rgb_input, random_shift, axis=1)   (kinetics_i3d_utils.py:115-116,135)
            clips = torch.roll(clips, int(self._rng.randint(0, self.frames)), dims=1).contiguous()
        if cycp:      # tf.roll(input_pert, random_shift_2, axis=0) (:130-131,137)
            shift_p = int(self._rng.randint(0, self.frames))
        a.beta0, a.beta1, a.beta2, a.beta3 = float(beta_0), float(beta_1), float(beta_2), float(beta_3)
        self._last_clips = clips
        if shift_p:
            self._step_rolled(clips, lab, flag, float(learning_rate), shift_p)
        else:
            a.step(clips, lab, adv_flag=flag, lr=float(learning_rate))
        sc = a.scalars.cpu().numpy()
        B = self.batch_size
        logits = a.eng.logits.cpu().numpy()
        probs = a.eng.probs.cpu().numpy()
        reg = beta_1 * sc[L.S_NORM_REG] + beta_2 * sc[L.S_DIFF_REG] + beta_3 * sc[L.S_LAP_REG]
        lab_np = lab.cpu().numpy()
        label_prob, max_non_label_prob, tmin, tmax = self._prob_handles(probs, lab_np)
        self._last = dict(
            loss=float(sc[L.S_ADV_LOSS] + beta_0 * reg), adversarial_loss=float(sc[L.S_ADV_LOSS]),
            adversarial_loss_total=float(sc[L.S_ADV_LOSS]), regularizer_loss=float(reg),
            norm_reg=float(sc[L.S_NORM_REG]), diff_norm_reg=float(sc[L.S_DIFF_REG]),
            laplacian_norm_reg=float(sc[L.S_LAP_REG]), thickness=float(sc[L.S_THICKNESS]),
            roughness=float(sc[L.S_ROUGHNESS]), thickness_relative=float(sc[L.S_THICKNESS]) / 2.0 * 100,
            roughness_relative=float(sc[L.S_ROUGHNESS]) / 2.0 * 100, to_min_prob=tmin, to_max_prob=tmax,
            model_logits=logits, softmax=probs, label_prob=label_prob, max_non_label_prob=max_non_label_prob,
            fooled_count=int(sc[L.S_FOOLED]))
        return self._last

    def _step_rolled(self, clips, lab, flag, lr, shift):
        """cyclic perturbation attack: the network sees roll(delta, shift); the gradient is rolled back."""
        self._atk.step_rolled(clips, lab, shift, adv_flag=flag, lr=lr)

    def adversarial_inputs_rgb_of(self, inputs):
        """sess.run(adversarial_inputs_rgb, {inputs: rgb_sample}) — fp32 [B,T,224,224,3]."""
        return self._atk.adversarial_video(self._to_device_clip(inputs)).cpu().numpy()

    def adversarial_video_uint8(self, inputs):
        """((adv+1.0)*127.5).astype(uint8) — utils/stats_and_plot/stats_plots.py:57, bit-exact."""
        return self._atk.adversarial_video(self._to_device_clip(inputs), as_uint8=True).cpu().numpy()

    def evaluate(self, next_element_val, targeted_attack=False, target_class_id=None, cyclic=0,
                 exclude_misclassify=True):
        """Fooling ratio over a validation iterable of (rgb_sample, sample_label) batches
        (utils/kinetics_i3d_utils.py:217-250).  Returns (miss_rate, total_val_vid)."""
        # Fused evaluation pass: the clean clips and the perturbed (possibly rolled) clips of a
        # batch go through ONE forward of the evaluation handle, the two counters stay on the device and are read once
        # after the last batch — the reference runs two sess.run per batch and counts on the host.
        self._configure()
        a = self._atk
        a.eval_counts(reset=True)
        for rgb_sample, sample_label in next_element_val:
            clips = self._to_device_clip(rgb_sample)
            clips_adv = None
            if cyclic:
                clips_adv = torch.roll(clips, int(self._rng.randint(0, self.frames)), dims=1).contiguous()
            lab = torch.as_tensor(np.asarray(sample_label), dtype=torch.int64).reshape(-1).to(self.device)
            a.eval_batch(clips, lab, clips_adv=clips_adv, targeted=targeted_attack, target_class=target_class_id,
                         exclude_misclassify=exclude_misclassify)
        miss, total = a.eval_counts(reset=True)
        # sharded validation: every rank evaluated its shard of the clips; the ratio is taken over all of them
        if a.world > 1:
            miss, total = fdist.sum_counts((miss, total), device=a.device, group=a.pg)
        return (miss / total if total else 0.0), int(total)


class kinetics_i3d_L12(_KineticsGraph):
    """Drop-in for ki3du.kinetics_i3d_L12(ckpt_path, batch_size, init_model, rgb_input, labels,
    cyclic_flag_default_c, default_adv_flag_c): the sparse per-pixel baseline (utils/kinetics_i3d_utils.py:308-521,
    FLICKERING_ATTACK=False).  eps_rgb is [T,224,224,3], initialised to 1e-8 (:333), not clipped (:336); the drivers
    minimise adversarial_loss + beta_1 * loss_L12 (i3d_adversarial_main_universal.py:133).  Same eager execution
    model as `kinetics_i3d` above."""

    flickering = False

    def __init__(self, ckpt_path="data/checkpoints/rgb_imagenet/model.ckpt", batch_size=1, init_model=True,
                 rgb_input=None, labels=None, cyclic_flag_default_c=0.0, default_adv_flag_c=1.0,
                 frames=_SAMPLE_VIDEO_FRAMES, weights=None, device=0, label_map_path=_LABEL_MAP_PATH, seed=0,
                 sharded=True):
        from .attack import SparseAttack
        w = self._init_common(ckpt_path, batch_size, frames, default_adv_flag_c, cyclic_flag_default_c, label_map_path,
                              weights, seed, rgb_input, labels)
        self._atk = SparseAttack(w, batch_size, frames, {}, device=device, sharded=sharded)
        self.device = self._atk.device

    def _to_device_clip(self, inputs):
        """The sparse engine path takes uint8 clips (its backward re-derives the range-clip mask from them).  The
        reference feeds float clips u8/128 - 1 (parse_example_uint8, utils/pre_process_rgb_flow.py:234): those are
        inverted exactly with (x + 1) * 128; floats that are not on that grid are refused instead of being silently
        re-quantised (round 1 inverted with 127.5 and moved most pixels by one level)."""
        t = torch.as_tensor(inputs)
        if t.dtype != torch.uint8:
            t = float_clip_to_u8(t.to(torch.float32), strict=True)
        t = t.reshape(self.batch_size, self.frames, _IMAGE_SIZE, _IMAGE_SIZE, 3)
        return t.to(self.device, non_blocking=True).contiguous()

    def train_step(self, inputs, labels, learning_rate=1e-3, beta_1=0.5, cyclic_flag=None, adv_flag=None):
        """One `sess.run([train_op, loss, adversarial_loss, regularizer_loss, thickness, roughness, ...])` with
        loss = adversarial_loss + beta_1 * loss_L12 (i3d_adversarial_main_universal.py:126-140)."""
        self._configure()
        a = self._atk
        clips = self._to_device_clip(inputs)
        lab = torch.as_tensor(np.asarray(labels), dtype=torch.int64).reshape(-1).to(self.device)
        cyc = self.cyclic_flag if cyclic_flag is None else float(cyclic_flag)
        if cyc:       # tf.roll(rgb_input, random_shift, axis=1)   (kinetics_i3d_utils.py:349-350,357)
            clips = torch.roll(clips, int(self._rng.randint(0, self.frames)), dims=1).contiguous()
        a.reg_weight = float(beta_1)
        a.step(clips, lab, adv_flag=self.adv_flag if adv_flag is None else float(adv_flag), lr=float(learning_rate))
        sc = a.scalars.cpu().numpy()
        B = self.batch_size
        logits = a.eng.logits.cpu().numpy()
        probs = a.eng.probs.cpu().numpy()
        lab_np = lab.cpu().numpy()
        label_prob, max_non_label_prob, tmin, tmax = self._prob_handles(probs, lab_np)
        l12 = float(sc[L.S_NORM_REG])            # fav_pixels_update: FAV_S_NORM_REG <- loss_L12
        self._last = dict(
            loss=float(sc[L.S_TOTAL_LOSS]), adversarial_loss=float(sc[L.S_ADV_LOSS]),
            adversarial_loss_total=float(sc[L.S_ADV_LOSS]), loss_L12=l12, regularizer_loss=float(beta_1) * l12,
            thickness=float(sc[L.S_THICKNESS]), roughness=float(sc[L.S_ROUGHNESS]),
            thickness_relative=float(sc[L.S_THICKNESS]) / 2.0 * 100,
            roughness_relative=float(sc[L.S_ROUGHNESS]) / 2.0 * 100, to_min_prob=tmin, to_max_prob=tmax,
            model_logits=logits, softmax=probs, label_prob=label_prob, max_non_label_prob=max_non_label_prob,
            fooled_count=int(sc[L.S_FOOLED]))
        return self._last

    def evaluate(self, next_element_val, targeted_attack=False, target_class_id=None, cyclic=0,
                 exclude_misclassify=True):
        """Fooling ratio over a validation iterable (utils/kinetics_i3d_utils.py:431-464)."""
        miss, total = 0, 0
        for rgb_sample, sample_label in next_element_val:
            clips = self._to_device_clip(rgb_sample)
            if cyclic:
                clips = torch.roll(clips, int(self._rng.randint(0, self.frames)), dims=1).contiguous()
            sample_label = np.asarray(sample_label).reshape(-1)
            pred = self._atk.predict(clips, adv_flag=1.0).cpu().numpy().argmax(-1)
            miss_cond = (pred == target_class_id) if targeted_attack else (pred != sample_label)
            if exclude_misclassify:
                clean = self._atk.predict(self._to_device_clip(rgb_sample), adv_flag=0.0).cpu().numpy().argmax(-1)
                valid = clean == sample_label
                miss += int(np.logical_and(miss_cond, valid).sum())
                total += int(valid.sum())
            else:
                miss += int(miss_cond.sum())
                total += int(miss_cond.shape[0])
        # sharded validation: every rank evaluated its shard of the clips; the ratio is taken over all of them
        if self._atk.world > 1:
            miss, total = fdist.sum_counts((miss, total), device=self._atk.device, group=self._atk.pg)
        return (miss / total if total else 0.0), int(total)
