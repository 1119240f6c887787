"""Attack loops of the three TF drivers, on top of the `kinetics_i3d` mirror.

  single_video_attack  — i3d_adversarial_main_single_video_npy.py:115-337  (SINGLE_VIDEO_ATTACK)
  class_gen_attack     — i3d_adversarial_main_single_class_gen.py:176-373  (CLASS_GEN_ATTACK)
  universal_attack     — i3d_adversarial_main_universal.py:45-203,353-380  (UNIVERSAL_ATTACK)

Differences kept on purpose: one forward + one backward per step instead of the
reference's three forwards; BATCH_SIZE and MAX_NUM_STEP from the YAML are honoured; data arrives as
iterables of (clip batch, labels) — TFRecord / .npy readers are SURVEY §8(f1).
"""
import os
import pickle

import numpy as np

from . import checkpoint as fckpt
from . import dist as fdist


def _attack_kwargs(cfg):
    return dict(learning_rate=0.001, beta_0=cfg.LAMBDA, beta_1=cfg.BETA_1, beta_2=cfg.BETA_2, beta_3=cfg.BETA_2,
                cyclic_flag=float(cfg.CYCLIC_ATTACK))


def _select_loss(k_i3d, cfg):
    if cfg.IMPROVE_ADV_LOSS:
        k_i3d.improve_adversarial_loss(margin=cfg.PROB_MARGIN, targeted=cfg.TARGETED_ATTACK, logits=cfg.USE_LOGITS)
    else:
        k_i3d.ce_adversarial_loss(targeted=cfg.TARGETED_ATTACK)


def single_video_attack(k_i3d, rgb_sample, correct_cls_id, cfg, result_path=None, max_extra_steps=None,
                        keep_history=True, log_every=0):
    """Attack one clip until `step > MAX_NUM_STEP and adversarial`
    (i3d_adversarial_main_single_video_npy.py:211-337).  Returns the result dict of App. D
    (and writes `{cls}_beta1_{b1}_th_{:.2f}%_rg_{:.2f}%.pkl` when result_path is given), or None when
    the clean clip is misclassified (the reference skips those: :137-139)."""
    classes = k_i3d.get_kinetics_classes()
    rgb_sample = np.asarray(rgb_sample)
    model_softmax = k_i3d(rgb_sample, adv_flag=0)
    top_id = int(model_softmax.argmax())
    if top_id != int(correct_cls_id):
        return None
    _select_loss(k_i3d, cfg)
    if cfg.TARGETED_ATTACK:
        target_class_id = classes.index(cfg.TARGETED_CLASS)
    else:
        target_class_id = int(correct_cls_id)
    kw = _attack_kwargs(cfg)
    res = {"correct_cls_prob": float(model_softmax.max()), "correct_cls": classes[int(correct_cls_id)],
           "correct_cls_id": int(correct_cls_id), "softmax_init": model_softmax, "rgb_sample": rgb_sample}
    hist = {k: [] for k in ("total_loss_l", "adv_loss_l", "reg_loss_l", "norm_reg_loss_l", "diff_norm_reg_loss_l",
                            "perturbation", "softmax", "fatness", "smoothness")}
    k_i3d.reset()
    step, max_step = 0, int(cfg.MAX_NUM_STEP)
    hard_stop = None if max_extra_steps is None else max_step + int(max_extra_steps)
    # the reference feeds the same host clip through feed_dict at every step; here it is uploaded once
    dev_clip = k_i3d._to_device_clip(rgb_sample) if hasattr(k_i3d, "_to_device_clip") else rgb_sample
    while True:
        out = k_i3d.train_step(dev_clip, [target_class_id], **kw)
        # adversarial test on the UPDATED perturbation (the reference's second sess.run, :217)
        soft = k_i3d(dev_clip, adv_flag=1) if (step > max_step or keep_history) else out["softmax"]
        pred = int(soft.argmax(-1)[0])
        is_adv = (pred == target_class_id) if cfg.TARGETED_ATTACK else (pred != target_class_id)
        if keep_history:
            hist["total_loss_l"].append(out["loss"])
            hist["adv_loss_l"].append(out["adversarial_loss"])
            hist["reg_loss_l"].append(out["regularizer_loss"])
            hist["norm_reg_loss_l"].append(out["norm_reg"])
            hist["diff_norm_reg_loss_l"].append(out["diff_norm_reg"])
            hist["perturbation"].append(k_i3d.eps_rgb.copy())
            hist["softmax"].append(soft)
            hist["fatness"].append(out["thickness"] / 2.0 * 100)
            hist["smoothness"].append(out["roughness"] / 2.0 * 100)
        if log_every and step % log_every == 0:
            print("Step: {:05d}, Total Loss: {:.5f}, Cls Loss: {:.5f}, Total Reg Loss: {:.5f}, thickness: {:.5f} "
                  "({:.2f} %), roughness: {:.5f} ({:.2f} %)".format(
                      step, out["loss"], out["adversarial_loss"], out["regularizer_loss"], out["thickness"],
                      out["thickness"] / 2.0 * 100, out["roughness"], out["roughness"] / 2.0 * 100))
        if (step > max_step and is_adv) or (hard_stop is not None and step >= hard_stop):
            if not keep_history:
                hist["perturbation"].append(k_i3d.eps_rgb.copy())
                hist["softmax"].append(soft)
                hist["fatness"].append(out["thickness"] / 2.0 * 100)
                hist["smoothness"].append(out["roughness"] / 2.0 * 100)
            res.update(hist)
            res["adv_video"] = k_i3d.adversarial_inputs_rgb_of(rgb_sample)
            res["total_steps"] = step
            res["beta_0"], res["beta_1"], res["beta_2"], res["beta_3"] = kw["beta_0"], kw["beta_1"], kw["beta_2"], kw["beta_3"]
            res["is_adversarial"] = bool(is_adv)
            if result_path:
                os.makedirs(result_path, exist_ok=True)
                fn = "{}_beta1_{}_th_{:.2f}%_rg_{:.2f}%.pkl".format(
                    classes[int(correct_cls_id)].replace(" ", "_"), kw["beta_1"], res["fatness"][-1], res["smoothness"][-1])
                with open(os.path.join(result_path, fn), "wb") as f:
                    pickle.dump(res, f)
                res["pkl_path"] = os.path.join(result_path, fn)
            return res
        step += 1


def record_batches_from_config(cfg, frames=None, rank=0, world=1, pinned=False, num_parallel_reads=None,
                               train_repeat=1):
    """(train_batches, val_batches) — the batch sources `class_gen_attack` / `universal_attack` take — from the
    record keys of a run_config.yml section, as the mains build them (i3d_adversarial_main_single_class_gen.py:111-135,
    i3d_adversarial_main_universal.py:205-245): `*.tfrecords` under every TF_RECORDS_{TRAIN,VAL}_PATH entry in sorted
    order, cut to NUM_OF_{TRAIN,VAL}_TF_RECORDS files when those keys exist, batches of BATCH_SIZE with the remainder
    dropped.  `num_parallel_reads` (the mains pass os.cpu_count()) selects tf.data's interleaved record order;
    `train_repeat`: the universal main repeats the training set 1000 times (:241).  With world > 1 every rank takes
    each world-th file (BATCH_SIZE is then the per-GPU batch)."""
    import glob
    from .records import ClipRecordDataset

    def files(paths, limit):
        paths = [paths] if isinstance(paths, str) else list(paths)
        out = []
        for p in paths:
            out += sorted(glob.glob(os.path.join(p, "*.tfrecords")))
        out = out[:int(limit)] if limit is not None else out
        return out[rank::world]

    train = files(cfg.TF_RECORDS_TRAIN_PATH, cfg.get("NUM_OF_TRAIN_TF_RECORDS"))
    val = files(cfg.TF_RECORDS_VAL_PATH, cfg.get("NUM_OF_VAL_TF_RECORDS"))
    if not train or not val:
        raise FileNotFoundError("no *.tfrecords under TF_RECORDS_TRAIN_PATH / TF_RECORDS_VAL_PATH for this rank")
    B = int(cfg.BATCH_SIZE)
    kw = dict(frames=frames, pinned=pinned, num_parallel_reads=num_parallel_reads)
    train_ds = ClipRecordDataset(train, B, repeat=train_repeat, **kw)

    def train_batches():
        return iter(train_ds)
    # ranks of a sharded run hold different files: the drivers cut every pass to the shortest shard (dist.lockstep)
    train_batches.num_batches = train_ds.num_batches
    return train_batches, (lambda: iter(ClipRecordDataset(val, B, **kw)))


def _train_iter(k_i3d, train_batches):
    """One pass over this rank's training batches; sharded runs stay in lockstep (`dist.lockstep`): every rank runs the
    number of steps the SHORTEST shard allows.  `train_batches.num_batches` (set by `record_batches_from_config`),
    when present, lets the ranks agree once per pass instead of once per step."""
    atk = k_i3d._atk
    if atk.world <= 1:
        return train_batches()
    n = getattr(train_batches, "num_batches", None)
    return fdist.lockstep(train_batches(), atk.world, atk.device, atk.pg, num_batches=n() if callable(n) else n)


def class_gen_attack(k_i3d, train_batches, val_batches, cfg, result_path=None, epochs=1, log_every=0, ckpt_prefix=None,
                     save_every=0):
    """One perturbation over batches of one class; fooling rate on the validation set after every
    pass; `res.pkl` layout of i3d_adversarial_main_single_class_gen.py:358-372.  `train_batches` /
    `val_batches` are callables returning a fresh iterator of (clips, labels) (the reference re-inits
    its tf.data iterators on OutOfRangeError, :334-337).  Stops at MAX_NUM_STEP.
    ckpt_prefix (the reference's `ckpt_dst`, :192-197,214,373): the run restores the newest
    `<ckpt_prefix>model_step_XXXXX.npz` (perturbation, Adam moments, step) before it starts, saves one at the start and
    after every pass over the training batches, and additionally every `save_every` steps when that is > 0."""
    classes = k_i3d.get_kinetics_classes()
    _select_loss(k_i3d, cfg)
    kw = _attack_kwargs(cfg)
    target_class_id = classes.index(cfg.TARGETED_CLASS) if cfg.TARGETED_ATTACK else None
    res = {k: [] for k in ("total_loss_l", "adv_loss_l", "reg_loss_l", "norm_reg_loss_l", "diff_norm_reg_loss_l",
                           "perturbation", "fatness", "smoothness", "fool_rate")}
    step, max_step = 0, int(cfg.MAX_NUM_STEP)
    if ckpt_prefix:
        step = fckpt.resume(ckpt_prefix, k_i3d._atk)
        fckpt.save_checkpoint(ckpt_prefix, k_i3d._atk, step)
    miss_rate, _ = k_i3d.evaluate(val_batches(), targeted_attack=cfg.TARGETED_ATTACK, target_class_id=target_class_id)
    res["fool_rate"].append(miss_rate)
    for _ in range(epochs):
        if step >= max_step:
            break
        for rgb_sample, sample_label in _train_iter(k_i3d, train_batches):
            labels = [target_class_id] * k_i3d.batch_size if cfg.TARGETED_ATTACK else sample_label
            out = k_i3d.train_step(rgb_sample, labels, **kw)
            res["total_loss_l"].append(out["loss"])
            res["adv_loss_l"].append(out["adversarial_loss"])
            res["reg_loss_l"].append(out["regularizer_loss"])
            res["norm_reg_loss_l"].append(out["norm_reg"])
            res["diff_norm_reg_loss_l"].append(out["diff_norm_reg"])
            res["fatness"].append(out["thickness"] / 2.0 * 100)
            res["smoothness"].append(out["roughness"] / 2.0 * 100)
            res["perturbation"].append(k_i3d.eps_rgb.copy())
            if log_every and step % log_every == 0:
                print("Step: {:05d}, Total Loss: {:.5f}, Cls Loss: {:.5f}".format(step, out["loss"], out["adversarial_loss"]))
            step += 1
            if ckpt_prefix and save_every and step % int(save_every) == 0:
                fckpt.save_checkpoint(ckpt_prefix, k_i3d._atk, step)
            if step >= max_step:
                break
        miss_rate, _ = k_i3d.evaluate(val_batches(), targeted_attack=cfg.TARGETED_ATTACK, target_class_id=target_class_id)
        res["fool_rate"].append(miss_rate)
        if ckpt_prefix:
            fckpt.save_checkpoint(ckpt_prefix, k_i3d._atk, step)
        res["total_steps"], res["beta_1"], res["beta_2"] = step, kw["beta_1"], kw["beta_2"]
        if result_path and fdist.is_writer(k_i3d._atk.world):
            os.makedirs(result_path, exist_ok=True)
            with open(os.path.join(result_path, "res.pkl"), "wb") as f:
                pickle.dump(res, f)
        if step >= max_step:
            break
    return res


def universal_attack(k_i3d, train_batches, val_batches, cfg, max_steps=None, eval_every=None, log_every=0,
                     summary_dir=None, model_dir=None, save_checkpoints_steps=100, keep_checkpoint_max=5):
    """UNIVERSAL_ATTACK with FLICKERING_ATTACK=True: same step as class-gen over all classes; returns
    {'perturbation': [T,1,1,3], 'fool_rate': [...], 'scalars': TensorBoard-tag -> list}
    (tags of i3d_adversarial_main_universal.py:176-196).
    model_dir: the estimator's checkpoint directory (universal.py:300-348) — the perturbation, the Adam moments and the
    global step are saved there every `save_checkpoints_steps` steps (newest `keep_checkpoint_max` kept, RunConfig :313-320)
    and a run started on a directory that holds a checkpoint continues from it (`tf.train.latest_checkpoint`, :334-348)."""
    sparse = not getattr(k_i3d, "flickering", True)     # FLICKERING_ATTACK=False: kinetics_i3d_L12, loss = adv + beta_0*beta_1*L12
    classes = k_i3d.get_kinetics_classes()
    _select_loss(k_i3d, cfg)
    if sparse:
        kw = dict(learning_rate=0.001, beta_1=float(cfg.LAMBDA) * float(cfg.BETA_1), cyclic_flag=float(cfg.CYCLIC_ATTACK))
    else:
        kw = _attack_kwargs(cfg)
        kw["cyclic_pert_flag"] = float(cfg.get("CYCLIC_PERTURBATION_ATTACK", False))
        kw["beta_3"] = cfg.BETA_2                      # universal.py:130 weights the laplacian term by beta_2
    target_class_id = classes.index(cfg.TARGETED_CLASS) if cfg.TARGETED_ATTACK else None
    max_steps = int(cfg.MAX_NUM_STEP if max_steps is None else max_steps)
    tags = {t: [] for t in ("Loss/total", "Loss/adversarial_loss", "Loss/regularizer_loss", "Loss/thickness",
                            "Loss/first_order_temporal_diff", "Loss/second_order_temporal_diff",
                            "Perturbation/thickness_%", "Perturbation/roughness_%", "Perturbation/max",
                            "Perturbation/min", "Probability/prob_to_min", "Probability/prob_to_max")}
    fool = []
    step = 0
    ckpt_prefix = os.path.join(model_dir, "") if model_dir else None
    if ckpt_prefix:
        step = fckpt.resume(ckpt_prefix, k_i3d._atk)
    writer = None
    if summary_dir and fdist.is_writer(k_i3d._atk.world):      # <model_dir>/train/events.out.tfevents.* like SummarySaverHook (universal.py:198-201); sharded runs: rank 0 writes
        from .records import SummaryWriter
        writer = SummaryWriter(os.path.join(summary_dir, "train"))
    while step < max_steps:
        n_before = step
        for rgb_sample, sample_label in _train_iter(k_i3d, train_batches):
            labels = [target_class_id] * k_i3d.batch_size if cfg.TARGETED_ATTACK else sample_label
            out = k_i3d.train_step(rgb_sample, labels, **kw)
            if step % 50 == 0:      # SummarySaverHook(save_steps=50) universal.py:198-201
                k_i3d._atk.check_replicas()       # sharded run: the replicated perturbation must be identical on all ranks
                eps = k_i3d.eps_rgb
                vals = (out["loss"], out["adversarial_loss"], out["regularizer_loss"], out.get("norm_reg", float("nan")),
                        out.get("diff_norm_reg", float("nan")), out.get("laplacian_norm_reg", float("nan")), out["thickness_relative"],
                        out["roughness_relative"], float(eps.max()), float(eps.min()),
                        float(np.mean(out["to_min_prob"])), float(np.mean(out["to_max_prob"])))
                for t, v in zip(tags, vals):
                    tags[t].append((step, v))
                    if writer is not None and v == v:
                        writer.add_scalar(t, v, step)
            if log_every and step % log_every == 0:
                print("step {:05d} loss {:.5f}".format(step, out["loss"]))
            step += 1
            if ckpt_prefix and save_checkpoints_steps and step % int(save_checkpoints_steps) == 0:
                fckpt.save_checkpoint(ckpt_prefix, k_i3d._atk, step, keep_max=keep_checkpoint_max)
            if eval_every and step % eval_every == 0:
                fool.append((step, k_i3d.evaluate(val_batches(), targeted_attack=cfg.TARGETED_ATTACK,
                                                  target_class_id=target_class_id)[0]))
            if step >= max_steps:
                break
        if step == n_before:
            raise RuntimeError("universal_attack: the training batch source is empty (on at least one rank)")
    fool.append((step, k_i3d.evaluate(val_batches(), targeted_attack=cfg.TARGETED_ATTACK,
                                      target_class_id=target_class_id)[0]))
    if ckpt_prefix:
        fckpt.save_checkpoint(ckpt_prefix, k_i3d._atk, step, keep_max=keep_checkpoint_max)
    if writer is not None:
        for st, fr in fool:        # eval metric 'ACC: 1- FOOLING_RATIO' (universal.py:160-161) is 1 - fool rate
            writer.add_scalar("ACC: 1- FOOLING_RATIO", 1.0 - fr, st)
        writer.close()
    return {"perturbation": k_i3d.eps_rgb, "fool_rate": fool, "scalars": tags, "total_steps": step}
