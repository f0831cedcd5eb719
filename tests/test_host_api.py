"""CPU-only checks of the host-side mirror of the reference interface: state_dict compatibility, registries,
error behaviour, sharding helper.  Where /root/reference is present the key sets are compared live."""
import pytest
import torch

from tests import synth
from oracle import ref_shim
from tests.util import torch_sd
from trphysx_b200 import AutoEmbeddingModel, AutoPhysConfig, AutoPhysTransformer, shard_bounds
from trphysx_b200.config import CylinderConfig, GrayScottConfig, LorenzConfig, PhysConfig, RosslerConfig
from trphysx_b200.embedding import CylinderEmbedding, GrayScottEmbedding, LorenzEmbedding, RosslerEmbedding
from trphysx_b200.transformer import PhysformerGPT2, PhysformerTrain
from trphysx_b200.transformer.phys_transformer_gpt2 import position_table


def test_state_dict_layouts_load_synth_weights():
    for kind, cfgc in (("lorenz", LorenzConfig), ("rossler", RosslerConfig), ("cylinder", CylinderConfig)):
        ns = synth.make_config(kind)
        m = PhysformerGPT2(cfgc(), kind)
        m.load_state_dict(torch_sd(synth.gpt2_weights(ns, 1)), strict=True)
        assert m.model_name == "transformer_" + kind
    LorenzEmbedding(LorenzConfig()).load_state_dict(torch_sd(synth.lorenz_embedding_weights(synth.make_config("lorenz"), 1)))
    RosslerEmbedding(RosslerConfig()).load_state_dict(torch_sd(synth.lorenz_embedding_weights(synth.make_config("rossler"), 1)))
    CylinderEmbedding(CylinderConfig()).load_state_dict(torch_sd(synth.cylinder_embedding_weights(synth.make_config("cylinder"), 1)))


def test_parameter_counts_match_reference_notebooks():
    # examples/lorenz/train_lorenz_enn.ipynb:437, examples/cylinder/train_cylinder_transformer.ipynb:479,481
    assert LorenzEmbedding(LorenzConfig()).num_parameters == 36192
    assert CylinderEmbedding(CylinderConfig()).num_parameters == 262535
    assert PhysformerGPT2(CylinderConfig(), "cylinder")._num_parameters() == 1208448
    assert PhysformerGPT2(LorenzConfig(), "lorenz")._num_parameters() == 53984


def _pe_like_forward(cfg):
    """phys_transformer_gpt2.py:215-219 evaluated the way forward() does it."""
    pos = torch.arange(0, cfg.n_ctx, dtype=torch.float).unsqueeze(0)
    pe = torch.zeros(1, cfg.n_ctx, cfg.n_embd)
    i = torch.arange(0, cfg.n_embd // 2, dtype=torch.float).unsqueeze(0).unsqueeze(0)
    pe[:, :, ::2] = torch.sin(pos.unsqueeze(-1) / 10000 ** (2 * i / cfg.n_embd))
    i = i[:, :, cfg.n_embd % 2]
    pe[:, :, 1::2] = torch.cos(pos.unsqueeze(-1) / 10000 ** (2 * i / cfg.n_embd))
    return pe[0]


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")
def test_state_dict_keys_equal_reference():
    ref_shim.load()
    from trphysx.config.configuration_lorenz import LorenzConfig as RL
    from trphysx.config.configuration_cylinder import CylinderConfig as RC
    from trphysx.embedding import LorenzEmbedding as RLE, CylinderEmbedding as RCE
    from trphysx.transformer import PhysformerGPT2 as RG

    def sig(m):
        return {k: (tuple(v.shape), v.dtype) for k, v in m.state_dict().items()}

    assert sig(PhysformerGPT2(LorenzConfig(), "x")) == sig(RG(RL(), "x"))
    assert sig(PhysformerGPT2(CylinderConfig(), "x")) == sig(RG(RC(), "x"))
    assert sig(LorenzEmbedding(LorenzConfig())) == sig(RLE(RL()))
    assert sig(CylinderEmbedding(CylinderConfig())) == sig(RCE(RC()))
    from trphysx.config.configuration_grayscott import GrayScottConfig as RGC
    from trphysx.embedding.embedding_grayscott import GrayScottEmbedding as RGE
    assert sig(GrayScottEmbedding(GrayScottConfig())) == sig(RGE(RGC()))
    assert list(GrayScottEmbedding(GrayScottConfig()).state_dict()) == list(RGE(RGC()).state_dict())
    assert vars(GrayScottConfig()).items() >= {k: v for k, v in vars(RGC()).items() if k in ("n_ctx", "n_embd", "n_layer", "n_head", "state_dims")}.items()
    # a reference config object is accepted as-is, and reference checkpoints load
    ours = PhysformerGPT2(RL(), "lorenz")
    ours.load_state_dict(RG(RL(), "lorenz").state_dict(), strict=True)
    assert torch.equal(position_table(64, 32), _pe_like_forward(RL()))


def test_registries_and_errors(tmp_path):
    assert isinstance(AutoPhysConfig.load_config("lorenz"), LorenzConfig)
    assert AutoPhysConfig.load_config("cylinder").n_embd == 128
    with pytest.raises(EnvironmentError):
        AutoEmbeddingModel()
    with pytest.raises(EnvironmentError):
        AutoPhysConfig()
    with pytest.raises(ValueError):
        AutoEmbeddingModel.init_model("nope", LorenzConfig())
    with pytest.raises(KeyError):
        AutoEmbeddingModel.init_trainer("nope", LorenzConfig())
    assert isinstance(AutoEmbeddingModel.init_model("lorenz", LorenzConfig()), LorenzEmbedding)
    gs = AutoEmbeddingModel.init_model("grayscott", AutoPhysConfig.load_config("grayscott"))
    assert isinstance(gs, GrayScottEmbedding) and gs.model_name == "embedding_grayscott" and gs.embed_dims == 512
    gs.load_state_dict(torch_sd(synth.grayscott_embedding_weights(synth.make_config("grayscott"), 3)), strict=True)
    with pytest.raises(NotImplementedError):
        AutoEmbeddingModel.init_trainer("grayscott", GrayScottConfig())(torch.zeros(1, 2, 2, 32, 32, 32))
    assert isinstance(AutoPhysTransformer.init_model("rossler"), PhysformerGPT2)
    m = AutoEmbeddingModel.init_model("lorenz", LorenzConfig())
    m.save_model(str(tmp_path), epoch=3)
    assert (tmp_path / "embedding_lorenz3.pth").exists()
    m2 = AutoEmbeddingModel.load_model("lorenz", LorenzConfig(), str(tmp_path), epoch=3)
    assert torch.equal(m2.kMatrixUT, m.kMatrixUT)
    with pytest.raises(FileNotFoundError):
        m.load_model(str(tmp_path / "missing"))
    t = PhysformerGPT2(LorenzConfig(), "lorenz")
    t.save_model(str(tmp_path), epoch=1)
    assert (tmp_path / "transformer_lorenz1.pth").exists()
    with pytest.raises(AssertionError):
        t.save_model(str(tmp_path / "transformer_lorenz1.pth"))
    cfg = LorenzConfig(n_ctx=16, foo=3)
    assert cfg.n_ctx == 16 and cfg.foo == 3 and cfg.hidden_size == 32
    cfg.save_pretrained(str(tmp_path / "cfg"))
    again = AutoPhysConfig.load_config(str(tmp_path / "cfg"))
    assert again.n_ctx == 16 and isinstance(again, LorenzConfig)
    with pytest.raises(KeyError):
        PhysConfig(n_ctx=1)


def test_cpu_tensors_are_refused_not_silently_computed():
    t = PhysformerGPT2(LorenzConfig(), "lorenz").eval()
    x = torch.zeros(2, 1, 32)
    with pytest.raises(NotImplementedError):
        t.generate(x, max_length=4)
    with pytest.raises(NotImplementedError):
        t(x)
    e = LorenzEmbedding(LorenzConfig()).eval()
    with pytest.raises(NotImplementedError):
        e.embed(torch.zeros(2, 3))
    with pytest.raises(NotImplementedError):
        e.recover(torch.zeros(2, 32))
    with pytest.raises(NotImplementedError):
        PhysformerGPT2(LorenzConfig(activation_function="relu"))
    head = PhysformerTrain(LorenzConfig(), t)
    assert head.transformer is t


def test_shard_bounds_partition():
    for n in (0, 1, 7, 8, 65536, 65537):
        for w in (1, 2, 3, 8):
            parts = [shard_bounds(n, w, r) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def test_blob_views_cover_the_packed_layout():
    """Gradient / master-parameter blob (csrc/gpt2_layout.cuh) <-> parameters in `_weight_list()` order: every view
    has its parameter's shape, the views tile the blob exactly, and mlp_f.weight is the transposed one."""
    for cfgc in (LorenzConfig, CylinderConfig):
        m = PhysformerGPT2(cfgc(), "x")
        params = m._weight_list()
        total = sum(p.numel() for p in params)
        E, nl = m.config.n_embd, m.config.n_layer
        assert total == nl * (12 * E * E + 13 * E) + E * E + 3 * E
        blob = torch.arange(total, dtype=torch.float32)
        views = m._blob_views(blob)
        assert len(views) == len(params)
        for p, v in zip(params, views):
            assert tuple(p.shape) == tuple(v.shape)
        assert views[-2].stride() == (1, E) and views[-2][3, 5] == blob[total - E - E * E + 5 * E + 3]
        assert all(v.is_contiguous() for v in views[:-2]) and views[-1].is_contiguous()


def test_training_entry_points_refuse_cpu():
    """The fused training step and the fused eval_states have no CPU path either."""
    from trphysx_b200.evaluation import eval_states
    from trphysx_b200.optim import FusedTransformerStep
    m = PhysformerGPT2(RosslerConfig(), "rossler")
    with pytest.raises(NotImplementedError):
        FusedTransformerStep(m)
    head = PhysformerTrain(RosslerConfig(), m).train()
    x = torch.zeros(2, 4, 32)
    with pytest.raises(NotImplementedError):
        head(inputs_embeds=x, labels_embeds=x)
    emb = RosslerEmbedding(RosslerConfig())
    with pytest.raises(NotImplementedError):
        eval_states(emb, x, torch.zeros(2, 4, 3))


def test_native_handles_are_not_copyable_state():
    """module copies must never share-and-free native handles.  deepcopy / pickle give the copy a fresh
    empty cache; shallow copies (copy.copy, nn.DataParallel replicas take __dict__.copy()) share ONE owner object
    whose finalizer is the only destroyer."""
    import copy
    import io
    from trphysx_b200 import _native as N
    destroyed = []
    real = N.HandleCache._destroy_all
    try:
        N.HandleCache._destroy_all = staticmethod(lambda entries, sym: (destroyed.append((sym, list(entries))), entries.clear()))
        for m, sym in ((PhysformerGPT2(LorenzConfig(), "x"), "trpx_gpt2_destroy"),
                       (LorenzEmbedding(LorenzConfig()), "trpx_mlpemb_destroy"),
                       (CylinderEmbedding(CylinderConfig()), "trpx_cyl_destroy")):
            assert isinstance(m._handles, N.HandleCache) and m._handles.destroy_symbol == sym
            assert not hasattr(type(m), "__del__")
            m._handles[0] = ("fake-handle", N.WeightTracker())         # pretend a kernel call created one
            deep = copy.deepcopy(m)
            assert deep._handles is not m._handles and len(deep._handles) == 0 and len(m._handles) == 1
            shallow = copy.copy(m)
            assert shallow._handles is m._handles
            buf = io.BytesIO()
            torch.save(m, buf)                                          # ctypes pointers used to make this raise
            buf.seek(0)
            loaded = torch.load(buf, weights_only=False)
            assert len(loaded._handles) == 0 and loaded._handles.destroy_symbol == sym
            m.invalidate_weights()
            assert m._handles[0][1].key is None
    finally:
        N.HandleCache._destroy_all = real


def test_feature_cache_format_matches_reference(tmp_path):
    """the pickle feature cache of trphysx/data_utils/dataset_phys.py:60-65,88-111 — same file name, same
    (examples, states) tuple of CPU tensors; a cache written here is read back by the reference's own read_cache (when
    the reference is importable) and one written with the reference's recipe is read here."""
    import pickle
    from trphysx_b200.data_utils import cached_features_file, read_cache, write_cache
    emb = LorenzEmbedding(LorenzConfig())
    data = tmp_path / "lorenz_valid.hdf5"
    data.write_bytes(b"")
    name = cached_features_file(str(data), emb, 64, ndata=8)
    assert name == str(tmp_path / "cached8_LorenzEmbedding_64_lorenz_valid.hdf5")
    assert cached_features_file(str(data), emb, 64, cache_path=str(tmp_path / "missing")) == \
        str(tmp_path / "cached-1_LorenzEmbedding_64_lorenz_valid.hdf5")
    examples = [torch.randn(64, 32) for _ in range(3)]
    states = [torch.randn(64, 3) for _ in range(3)]
    write_cache(name, examples, states)
    with open(name, "rb") as f:                                   # exactly what dataset_phys.py:97 unpickles
        ex2, st2 = pickle.load(f)
    assert isinstance(ex2, list) and all(torch.equal(a, b) for a, b in zip(ex2, examples))
    assert all(torch.equal(a, b) for a, b in zip(st2, states))
    ref_style = tmp_path / "ref_cache"
    with open(ref_style, "wb") as f:                              # dataset_phys.py:107-108
        pickle.dump((examples, []), f, protocol=pickle.HIGHEST_PROTOCOL)
    ex3, st3 = read_cache(str(ref_style))
    assert len(ex3) == 3 and st3 == [] and torch.equal(ex3[1], examples[1])
    if ref_shim.available():
        ref_shim.load()
        from trphysx.data_utils.dataset_phys import PhysicalDataset

        class _Probe(PhysicalDataset):                            # read_cache without the HDF5 constructor
            def __init__(self):
                pass

            def embed_data(self, *a, **k):
                pass
        p = _Probe()
        p.read_cache(name)
        assert torch.equal(p.examples[2], examples[2]) and torch.equal(p.states[0], states[0])
