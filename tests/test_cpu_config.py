"""not gpu: run_config.yml surface (keys verbatim from the reference's run_config.yml)."""
import yaml

from flickering_adversarial_video_b200 import config


def test_default_config_has_every_reference_key(tmp_path):
    cfg = config.default_config()
    assert config.validate(cfg) == []
    assert cfg.SINGLE_VIDEO_ATTACK.MAX_NUM_STEP == 2500 and cfg.CLASS_GEN_ATTACK.LAMBDA == 10.0
    assert cfg.UNIVERSAL_ATTACK.FLICKERING_ATTACK is True and cfg.UNIVERSAL_ATTACK.BATCH_SIZE == 8
    p = tmp_path / "run_config.yml"
    p.write_text(yaml.safe_dump({k: dict(v) for k, v in config.DEFAULTS.items()}))
    loaded = config.load_config(str(p))
    assert loaded.MODEL.CKPT_PATH == cfg.MODEL.CKPT_PATH
    assert loaded.UNIVERSAL_ATTACK.TF_RECORDS_TRAIN_PATH == cfg.UNIVERSAL_ATTACK.TF_RECORDS_TRAIN_PATH
    loaded.SINGLE_VIDEO_ATTACK.BETA_1 = 0.1          # attribute writes like easydict
    assert loaded["SINGLE_VIDEO_ATTACK"]["BETA_1"] == 0.1


def test_missing_keys_are_reported(tmp_path):
    p = tmp_path / "c.yml"
    p.write_text("DATA: {LABEL_MAP_PATH: x}\n")
    missing = config.validate(config.load_config(str(p)))
    assert "MODEL" in missing and "UNIVERSAL_ATTACK" in missing
