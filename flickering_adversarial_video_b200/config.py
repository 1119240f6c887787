"""run_config.yml loader with the reference's keys, verbatim (run_config.yml:2-89), exposed as an
attribute dictionary like `edict(yaml.load(f))` in utils/kinetics_i3d_utils.py:22-26."""
import copy

import yaml


class AttrDict(dict):
    """dict with attribute access, recursive (easydict stand-in; the reference imports easydict)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {}, **kw)
        for k, v in d.items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            v = AttrDict(v)
        elif isinstance(v, (list, tuple)):
            v = type(v)(AttrDict(x) if isinstance(x, dict) else x for x in v)
        super().__setitem__(k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    __setattr__ = __setitem__

    def __deepcopy__(self, memo):
        return AttrDict(copy.deepcopy(dict(self), memo))


# Defaults == the values shipped in the reference's run_config.yml
_ATTACK_COMMON = dict(
    TARGETED_ATTACK=False, IMPROVE_ADV_LOSS=True, PROB_MARGIN=0.05, USE_LOGITS=False,
    BETA_1=0.5, BETA_2=0.5, CYCLIC_ATTACK=False, NPY_PATH="data/videos_for_tests/npy/")
DEFAULTS = dict(
    DATA=dict(LABEL_MAP_PATH="data/label_map.txt"),
    MODEL=dict(CKPT_PATH="data/checkpoints/rgb_imagenet/model.ckpt",
               CKPT_PATH_WITH_ZERO_PERT="data/checkpoints/rgb_imagenet_with_zero_pert/model_step_00000"),
    SINGLE_VIDEO_ATTACK=dict(
        _ATTACK_COMMON, TARGETED_CLASS="javelin throw", MAX_NUM_STEP=2500, LAMBDA=1.0, BATCH_SIZE=1,
        PKL_RESULT_PATH="result/videos_for_tests/npy/",
        TF_RECORDS_TRAIN_PATH="data/kinetics/database/tfrecord_uint8/val/",
        TF_RECORDS_VAL_PATH="data/kinetics/database/tfrecord_uint8/val/"),
    CLASS_GEN_ATTACK=dict(
        _ATTACK_COMMON, TARGETED_CLASS="javelin throw", MAX_NUM_STEP=10000, LAMBDA=10.0, BATCH_SIZE=8,
        PKL_RESULT_PATH="result/generalization/model_gen_one_class/",
        TF_RECORDS_TRAIN_PATH=["data/kinetics/database/tfrecord/test/hula hooping"],
        TF_RECORDS_VAL_PATH=["data/kinetics/database/tfrecord/test/hula hooping"],
        NUM_OF_TRAIN_TF_RECORDS=10, NUM_OF_VAL_TF_RECORDS=5, NUM_OF_VID_EACH_TF_RECORDS=100),
    UNIVERSAL_ATTACK=dict(
        _ATTACK_COMMON, FLICKERING_ATTACK=True, TARGETED_CLASS="welding", MAX_NUM_STEP=10000, LAMBDA=1.0,
        BATCH_SIZE=8, CYCLIC_PERTURBATION_ATTACK=False,
        PKL_RESULT_PATH="result/generalization/universal_untargeted/",
        TF_RECORDS_TRAIN_PATH=["data/kinetics/database/tfrecord/test_all_cls/"],
        TF_RECORDS_VAL_PATH=["data/kinetics/database/tfrecord/test_all_cls/"],
        NUM_OF_TRAIN_TF_RECORDS=21, NUM_OF_VAL_TF_RECORDS=40, NUM_OF_VID_EACH_TF_RECORDS=50),
)
SECTIONS = ("SINGLE_VIDEO_ATTACK", "CLASS_GEN_ATTACK", "UNIVERSAL_ATTACK")


def load_config(yml_path):
    """Same call as ki3du.load_config(yml_path='run_config.yml') (utils/kinetics_i3d_utils.py:22-26)."""
    with open(yml_path, "r") as f:
        cfg = yaml.safe_load(f)
    if not isinstance(cfg, dict):
        raise ValueError(f"{yml_path}: expected a mapping at top level")
    return AttrDict(cfg)


def default_config():
    return AttrDict(copy.deepcopy(DEFAULTS))


def validate(cfg):
    """Check that every key the reference drivers read is present; returns the list of missing keys."""
    missing = []
    for sec, keys in DEFAULTS.items():
        if sec not in cfg:
            missing.append(sec)
            continue
        for k in keys:
            if k not in cfg[sec]:
                missing.append(f"{sec}.{k}")
    return missing
