"""On-disk formats -> engine inputs: files are synthesised in the reference's formats (Lightning
state_dict with training-only entries, tokenizer.pth with a config object, SMPL pickle with chumpy arrays and a scipy
sparse regressor written by stand-in classes under the real module names) and read back WITHOUT those modules."""
import pickle
import sys
import types

import numpy as np
import pytest
import torch

from tokenhmr_b200 import checkpoint as C
from tokenhmr_b200 import synth
from tokenhmr_b200.config import tiny_config


def _fake_module(name, **classes):
    m = types.ModuleType(name)
    for k, v in classes.items():
        v.__module__ = name
        setattr(m, k, v)
    return m


class Ch:
    """Writer-side stand-in for chumpy.ch.Ch (pickled under that module name; the module is removed before loading)."""

    def __init__(self, x):
        self.x = np.asarray(x)

    def __getstate__(self):
        return {"x": self.x, "_dirty_vars": set(), "_itr": None}


Ch.__qualname__ = "Ch"


def _write_smpl_pkl(tmp_path, smpl, V):
    """A pickle that looks like SMPL_NEUTRAL.pkl: chumpy.ch.Ch objects (state {'x': array}) and a
    scipy.sparse csc_matrix J_regressor; the writer-side classes are registered only while pickling."""
    real_scipy = {k: sys.modules.get(k) for k in ("chumpy", "chumpy.ch")}
    sys.modules["chumpy"] = _fake_module("chumpy")
    sys.modules["chumpy.ch"] = _fake_module("chumpy.ch", Ch=Ch)
    import scipy.sparse as sp
    posedirs = smpl["posedirs"].numpy().T.reshape(V, 3, -1)            # (V,3,207) as stored on disk
    kin = np.stack([np.array([2 ** 32 - 1] + smpl["parents"].tolist()[1:], dtype=np.uint32),
                    np.arange(24, dtype=np.uint32)])
    data = {
        "v_template": smpl["v_template"].numpy().astype(np.float64),
        "shapedirs": Ch(np.concatenate([smpl["shapedirs"].numpy(), np.zeros((V, 3, 290))], -1).astype(np.float64)),
        "posedirs": posedirs.astype(np.float64),
        "J_regressor": sp.csc_matrix(smpl["J_regressor"].numpy().astype(np.float64)),
        "weights": smpl["lbs_weights"].numpy().astype(np.float64),
        "kintree_table": kin,
        "f": np.arange(30, dtype=np.uint32).reshape(10, 3),
        "J": Ch(np.zeros((24, 3))),
        "bs_style": "lbs",
    }
    p = tmp_path / "SMPL_NEUTRAL.pkl"
    with open(p, "wb") as f:
        pickle.dump(data, f, protocol=2)
    pj = tmp_path / "SMPL_to_J19.pkl"
    with open(pj, "wb") as f:
        pickle.dump(smpl["joint_regressor_extra"].numpy().astype(np.float64), f, protocol=2)
    for k, v in real_scipy.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v
    return p, pj


def test_smpl_pkl_is_read_without_chumpy(tmp_path):
    cfg = tiny_config()
    smpl = synth.make_smpl(cfg)
    V = cfg.num_verts
    p, pj = _write_smpl_pkl(tmp_path, smpl, V)
    assert "chumpy" not in sys.modules
    got = C.load_smpl_pkl(p, str(pj), num_betas=10)
    for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "lbs_weights", "joint_regressor_extra"):
        assert got[k].dtype == torch.float32 and got[k].shape == smpl[k].shape, k
        torch.testing.assert_close(got[k], smpl[k], rtol=0, atol=0)
    assert got["parents"].tolist() == smpl["parents"].tolist() and got["parents"][0] == -1
    assert got["faces"].shape == (10, 3) and got["extra_vertex_ids"].shape == (21,)
    # smplx's directory convention
    got2 = C.load_smpl_pkl(tmp_path, num_betas=5)
    assert got2["shapedirs"].shape == (V, 3, 5) and "joint_regressor_extra" not in got2


def test_sparse_stub_rebuilds_csr_csc_coo(tmp_path, monkeypatch):
    """Force the 'scipy missing' path: the pickled sparse state is densified by the stand-in."""
    import scipy.sparse as sp
    rng = np.random.default_rng(0)
    dense = rng.random((24, 431)) * (rng.random((24, 431)) < 0.05)
    for make in (sp.csc_matrix, sp.csr_matrix, sp.coo_matrix):
        blob = pickle.dumps(make(dense), protocol=2)
        real = C._TolerantUnpickler.find_class

        def find_class(self, module, name, _real=real):
            if module.startswith("scipy.sparse"):
                return C._stub_class(module, name)
            return _real(self, module, name)

        monkeypatch.setattr(C._TolerantUnpickler, "find_class", find_class)
        import io
        obj = C._TolerantUnpickler(io.BytesIO(blob), encoding="latin1").load()
        monkeypatch.setattr(C._TolerantUnpickler, "find_class", real)
        assert isinstance(obj, C._Stub)
        np.testing.assert_array_equal(C._to_array(obj), dense)


def test_lightning_and_tokenizer_checkpoints(tmp_path):
    cfg = tiny_config()
    sd = synth.make_state_dict(cfg)
    model_sd = {k: v for k, v in sd.items() if not k.startswith("tokenizer.")}
    extra = {"discriminator.fc.weight": torch.zeros(3, 3), "smpl_parameter_loss.w": torch.zeros(1)}
    torch.save({"state_dict": {**model_sd, **extra}, "epoch": 3, "optimizer_states": [{}]}, tmp_path / "tokenhmr_model.ckpt")
    got = C.load_lightning_state_dict(tmp_path / "tokenhmr_model.ckpt")
    assert set(got) == set(model_sd)

    # tokenizer.pth: hparams is an object of a class that does not exist at load time
    CfgNode = _CfgNode
    sys.modules["yacs_fake_for_test"] = _fake_module("yacs_fake_for_test", CfgNode=CfgNode)
    arch = CfgNode(CODE_DIM=cfg.code_dim, NB_CODE=cfg.nb_code, WIDTH=cfg.tok_width, DEPTH=cfg.tok_depth,
                   DILATION_RATE=cfg.tok_dilation_rate, TOKEN_SIZE_DIV=cfg.tok_size_div, TOKEN_SIZE_MUL=4, DOWN_T=1)
    hp = CfgNode(ARCH=arch, EXP_NAME="x")
    net = {k[len("tokenizer."):]: v for k, v in sd.items() if k.startswith("tokenizer.")}
    net["decoder.body_model.shapedirs"] = torch.zeros(4)
    net.update({k[len("tokenizer."):]: v for k, v in synth.make_tokenizer_encoder_state_dict(cfg).items()})
    torch.save({"net": net, "hparams": hp}, tmp_path / "tokenizer.pth")
    del sys.modules["yacs_fake_for_test"]
    net2, arch2 = C.load_tokenizer_checkpoint(tmp_path / "tokenizer.pth")
    assert not any("body_model" in k for k in net2)
    assert arch2["CODE_DIM"] == cfg.code_dim and arch2["TOKEN_SIZE_MUL"] == 4
    merged = C.merge_state_dicts(got, net2)
    for k, v in sd.items():
        assert torch.equal(merged[k], v), k
    assert "tokenizer.encoder.encoder.0.weight" in merged

    (tmp_path / "model_config.yaml").write_text(
        "MODEL:\n  IMAGE_SIZE: 256\n  BBOX_SHAPE: [192, 256]\n  SMPL_HEAD:\n    TYPE: token\n    TOKENIZER: {TOKEN_CODE_DIM: 256, TOKEN_NUM: 160, "
        "TOKEN_CLASS_NUM: 2048}\n    TRANSFORMER_DECODER: {depth: 6, heads: 8, mlp_dim: 1024, dim_head: 64}\nEXTRA:\n  FOCAL_LENGTH: 5000\n")
    c2 = C.config_from_files(str(tmp_path / "model_config.yaml"), arch2)
    from tokenhmr_b200.config import release_config
    assert c2 == release_config()


class _CfgNode(dict):
    pass


_CfgNode.__qualname__ = "CfgNode"
_CfgNode.__name__ = "CfgNode"


@pytest.mark.gpu
def test_load_tokenhmr_from_files_runs(tmp_path, cuda_dev):
    """The file-based load_tokenhmr builds an engine whose forward equals the one built from the in-memory dicts."""
    from tokenhmr_b200.engine import TokenHMREngine
    cfg = tiny_config(vit_depth=1)
    sd, smpl = synth.make_state_dict(cfg), synth.make_smpl(cfg)
    (tmp_path / "checkpoints").mkdir()
    (tmp_path / "body_models" / "smpl").mkdir(parents=True)
    torch.save({"state_dict": {k: v for k, v in sd.items() if not k.startswith("tokenizer.")}}, tmp_path / "checkpoints" / "m.ckpt")
    torch.save({"net": {k[len("tokenizer."):]: v for k, v in sd.items() if k.startswith("tokenizer.")},
                "hparams": {"ARCH": {"CODE_DIM": cfg.code_dim}}}, tmp_path / "checkpoints" / "tokenizer.pth")
    p, pj = _write_smpl_pkl(tmp_path / "body_models" / "smpl", smpl, cfg.num_verts)
    pj.rename(tmp_path / "body_models" / "SMPL_to_J19.pkl")
    model, cfg2 = C.load_tokenhmr(str(tmp_path / "checkpoints" / "m.ckpt"), device="cuda:0", use_cuda_graph=False)
    assert cfg2 == cfg                       # depth and mesh size are inferred from the files
    img = synth.make_images(2, cfg)
    ref = TokenHMREngine(cfg, sd, smpl, device=cuda_dev, use_cuda_graph=False)({"img": img})
    out = model({"img": img})
    for k in ("pred_vertices", "pred_keypoints_3d", "pred_cam", "cls_logits_softmax"):
        assert torch.equal(out[k], ref[k]), k


# ---------------------------------------------------------------------------------------------------------
# Variants of the reference the engine does not implement must be rejected, not silently run as the release path
_BASE_YAML = {
    "MODEL": {"IMAGE_SIZE": 256, "BBOX_SHAPE": [192, 256], "BACKBONE": {"TYPE": "vit"},
              "SMPL_HEAD": {"TYPE": "token", "TOKENIZER": {"TOKEN_CODE_DIM": 256, "TOKEN_NUM": 160, "TOKEN_CLASS_NUM": 2048,
                                                          "TOKENIZER_TYPE": "Vanilla"},
                            "TRANSFORMER_DECODER": {"depth": 6, "heads": 8, "mlp_dim": 1024, "dim_head": 64, "norm": "layer",
                                                    "context_dim": 1280}}},
    "SMPL": {"NUM_BODY_JOINTS": 23, "GENDER": "neutral"}, "EXTRA": {"FOCAL_LENGTH": 5000},
}


def _yaml_with(tmp_path, path, value):
    import copy
    import yaml
    y = copy.deepcopy(_BASE_YAML)
    node = y
    for k in path[:-1]:
        node = node.setdefault(k, {})
    node[path[-1]] = value
    p = tmp_path / ("cfg_" + "_".join(path) + ".yaml")
    p.write_text(yaml.safe_dump(y))
    return str(p)


def test_release_yaml_is_accepted(tmp_path):
    import yaml
    from tokenhmr_b200.config import release_config
    p = tmp_path / "ok.yaml"
    p.write_text(yaml.safe_dump(_BASE_YAML))
    assert C.config_from_files(str(p), {"NB_CODE": 2048, "CODE_DIM": 256}) == release_config()


@pytest.mark.parametrize("path,value", [
    (("MODEL", "SMPL_HEAD", "TYPE"), "transformer_decoder"),        # heads/__init__.py:4-13
    (("MODEL", "SMPL_HEAD", "IEF_ITERS"), 3),                       # token_head.py:86
    (("MODEL", "SMPL_HEAD", "TRANSFORMER_INPUT"), "mean_shape"),    # token_head.py:29,88-91
    (("MODEL", "SMPL_HEAD", "JOINT_REP"), "aa"),                    # token_head.py:23-24
    (("MODEL", "SMPL_HEAD", "TOKENIZER", "TOKENIZER_TYPE"), "parts"),
    (("MODEL", "SMPL_HEAD", "TRANSFORMER_DECODER", "norm"), "batch"),
    (("MODEL", "SMPL_HEAD", "TRANSFORMER_DECODER", "dim_head"), 32),
    (("MODEL", "BACKBONE", "TYPE"), "resnet"),
    (("SMPL", "update_hips"), True),                                 # smpl_wrapper.py:33-36
    (("SMPL", "NUM_BODY_JOINTS"), 21),
])
def test_unsupported_reference_variants_are_rejected(tmp_path, path, value):
    from tokenhmr_b200._lib import ThmrError
    with pytest.raises(ThmrError, match="unsupported configuration"):
        C.config_from_files(_yaml_with(tmp_path, path, value), None)


def test_weights_that_contradict_the_config_are_rejected():
    import dataclasses
    from tokenhmr_b200._lib import ThmrError
    cfg = tiny_config()
    sd, smpl = synth.make_state_dict(cfg), synth.make_smpl(cfg)
    C.validate_against_weights(cfg, sd, smpl)                                      # consistent: passes
    with pytest.raises(ThmrError, match="NB_CODE"):
        C.validate_against_weights(dataclasses.replace(cfg, nb_code=1024), sd, smpl)
    bad = dict(sd)
    bad["tokenizer.quantizer.codebook"] = torch.zeros(1024, cfg.code_dim)
    with pytest.raises(ThmrError, match="codebook"):
        C.validate_against_weights(cfg, bad, smpl)
    bad = dict(sd)
    bad["smpl_head.transformer.to_token_embedding.weight"] = torch.zeros(1024, 157)    # 'mean_shape' checkpoint
    with pytest.raises(ThmrError, match="mean_shape"):
        C.validate_against_weights(cfg, bad, smpl)
    no_extra = {k: v for k, v in smpl.items() if k != "joint_regressor_extra"}
    with pytest.warns(UserWarning, match="25 instead of 44"):
        C.validate_against_weights(cfg, sd, no_extra)


@pytest.mark.gpu
def test_engine_device_moves_outputs_and_batch_limit(cuda_dev):
    """`.to()` of another device raises (it used to be a silent no-op), forward returns fresh tensors by default (the
    reference's behaviour), and max_batch is enforced."""
    from tokenhmr_b200._lib import ThmrError
    from tokenhmr_b200.engine import TokenHMREngine
    cfg = tiny_config(vit_depth=1)
    model = TokenHMREngine(cfg, synth.make_state_dict(cfg), synth.make_smpl(cfg), device=cuda_dev, use_cuda_graph=False,
                           max_batch=4, max_cached_shapes=2)
    assert model.to(cuda_dev) is model and model.to("cuda:0") is model and model.eval() is model
    assert model.to(torch.float32) is model
    for bad in ("cpu", torch.device("cpu"), torch.float16):
        with pytest.raises(ThmrError):
            model.to(bad)
    with pytest.raises(ThmrError):
        model.cpu()
    with pytest.raises(ThmrError):
        model.half()
    a = model({"img": synth.make_images(2, cfg, seed=1)})
    keep = a["pred_vertices"]
    snapshot = keep.clone()
    b = model({"img": synth.make_images(2, cfg, seed=2)})
    assert torch.equal(keep, snapshot) and not torch.equal(b["pred_vertices"], keep)       # no aliasing by default
    c1 = model({"img": synth.make_images(2, cfg, seed=1)}, alias_outputs=True)["pred_vertices"]
    assert torch.equal(c1, snapshot)
    model({"img": synth.make_images(2, cfg, seed=2)}, alias_outputs=True)
    assert not torch.equal(c1, snapshot)                                                    # opt-in views do alias
    with pytest.raises(ThmrError, match="max_batch"):
        model({"img": synth.make_images(5, cfg)})
    for B in (1, 2, 3, 4, 1):                                # more shapes than max_cached_shapes: LRU eviction, still correct
        out = model({"img": synth.make_images(B, cfg, seed=1)})
        assert out["pred_vertices"].shape[0] == B
    assert len(model._bufs) <= 2
    assert torch.equal(model({"img": synth.make_images(2, cfg, seed=1)})["pred_vertices"], snapshot)
