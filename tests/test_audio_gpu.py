"""Audio front-end on CUDA: the wav2vec2-base forward against `transformers.Wav2Vec2Model` itself (the
third-party class the reference calls, Net.py:611-612/644, executed on the CPU in fp32 with the same random-init weights),
the per-frame token windows, the SpeedEncoder against vectors recorded from the reference class, and the tokens fed through
the pipeline's audio cross-attention path (BASELINE config #3 wiring)."""
from pathlib import Path

import pytest
import torch

from util_models import TINY_CFG, check_parity, rel_l2, rerandomise_zero_inits

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def w2v():
    from emote_hack_b200.audio import Wav2Vec2Model
    from oracle import ref_audio
    hf = ref_audio.hf_wav2vec2(seed=0)
    ours = Wav2Vec2Model()
    missing, unexpected = ours.load_state_dict(hf.state_dict(), strict=True)
    assert not missing and not unexpected
    return hf, ours.cuda().eval()


def _wave(seconds, seed):
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(int(16000 * seconds)) / 16000.0
    # speech-like: a few harmonics with a slow envelope + noise, not zero mean, not unit variance
    x = 0.3 * torch.sin(2 * torch.pi * 140 * t) * (0.6 + 0.4 * torch.sin(2 * torch.pi * 3 * t)) \
        + 0.1 * torch.sin(2 * torch.pi * 2250 * t) + 0.05 * torch.randn(t.numel(), generator=g) + 0.02
    return x


@pytest.mark.parametrize("seconds,seed", [(1.0, 1), (2.37, 2)])
def test_wav2vec2_forward_matches_transformers(w2v, seconds, seed):
    from oracle import ref_audio
    hf, ours = w2v
    x = _wave(seconds, seed)
    with torch.no_grad():
        ref = hf(ref_audio.normalize_waveform(x)[None]).last_hidden_state
    out = ours(x.cuda(), normalize=True).last_hidden_state
    assert out.shape == ref.shape == (1, ours.frames_for(x.numel()), 768) and out.dtype == torch.float32
    check_parity(f"audio.wav2vec2_last_hidden_{seconds}s", rel_l2(out, ref), 2e-2)
    # already-normalised input_values (the processor's output), [1, n] layout
    out2 = ours(ref_audio.normalize_waveform(x)[None].cuda()).last_hidden_state
    check_parity(f"audio.wav2vec2_prenormalised_input_{seconds}s", rel_l2(out2, ref), 2e-2)


def test_wav2vec2_legacy_weight_norm_keys_and_weight_updates(w2v):
    from emote_hack_b200.audio import Wav2Vec2Model
    hf, ours = w2v
    sd = dict(hf.state_dict())
    sd["encoder.pos_conv_embed.conv.weight_g"] = sd.pop("encoder.pos_conv_embed.conv.parametrizations.weight.original0")
    sd["encoder.pos_conv_embed.conv.weight_v"] = sd.pop("encoder.pos_conv_embed.conv.parametrizations.weight.original1")
    m = Wav2Vec2Model()
    m.load_state_dict(sd, strict=True)                       # the checkpoint spelling of facebook/wav2vec2-base-960h
    m = m.cuda()
    x = _wave(0.5, 3).cuda()
    a = m(x, normalize=True).last_hidden_state
    assert rel_l2(a, ours(x, normalize=True).last_hidden_state) < 1e-5
    with torch.no_grad():
        m.encoder.layers[3].feed_forward.output_dense.bias.add_(0.5)
    assert rel_l2(m(x, normalize=True).last_hidden_state, a) > 1e-4      # packed copies follow in-place updates
    with pytest.raises(ValueError):
        m(torch.zeros(300, device="cuda"))


def test_feature_extractor_windows_and_pipeline_audio_context(w2v):
    """Wav2VecFeatureExtractor (Net.py:607-667) -> [T, 5, 768] tokens -> `encoder_hidden_states` of the audio
    cross-attention through EMOAnimationPipeline.__call__(audio=...)."""
    from emote_hack_b200.audio import Wav2VecFeatureExtractor, window_features
    from emote_hack_b200.pipeline import DDIMScheduler, EMOAnimationPipeline
    from emote_hack_b200.unet3d import UNet3DConditionModel
    from emote_hack_b200.vae import AutoencoderKL
    from oracle import ref_audio
    hf, ours = w2v
    fx = Wav2VecFeatureExtractor(model=ours)
    x = _wave(0.3, 4)
    tok = fx.extract_tokens(x, m=2, n=2)
    T = ours.frames_for(x.numel())
    assert tok.shape == (T, 5, 768)
    with torch.no_grad():
        hid = hf(ref_audio.normalize_waveform(x)[None]).last_hidden_state[0]
    check_parity("audio.tokens_windows", rel_l2(tok, window_features(hid, 2, 2)), 2e-2)
    assert tok[0, :2].abs().sum() == 0 and tok[-1, -2:].abs().sum() == 0          # zero padding past both ends
    # tokens as the per-frame context of a (tiny, 768-wide context) video UNet through the reference call surface
    cfg = dict(TINY_CFG, cross_attention_dim=768)
    torch.manual_seed(0)
    unet = rerandomise_zero_inits(UNet3DConditionModel(**cfg).eval()).cuda()
    torch.manual_seed(1)
    vae = AutoencoderKL(block_out_channels=(64, 64, 128, 128)).eval().cuda()
    pipe = EMOAnimationPipeline(vae, unet, DDIMScheduler(), audio_encoder=fx)
    emb = torch.randn(2, 7, 768, generator=torch.Generator().manual_seed(2)).cuda()
    frames = min(T, 8)
    lat = torch.randn(1, 4, frames, 8, 8, generator=torch.Generator().manual_seed(3)).cuda()
    a = pipe(emb, video_length=frames, height=64, width=64, num_inference_steps=2, latents=lat.clone(), audio=x).videos
    b = pipe(emb, video_length=frames, height=64, width=64, num_inference_steps=2, latents=lat.clone(), audio_features=tok).videos
    c = pipe(emb, video_length=frames, height=64, width=64, num_inference_steps=2, latents=lat.clone()).videos
    assert a.shape == (1, 3, frames, 64, 64) and torch.equal(a, b)
    assert (a - c).abs().max() > 1e-3                      # the audio tokens condition the result


def test_speed_encoder_matches_reference_vectors():
    """SpeedEncoder (Net.py:198-258) against outputs recorded by executing the reference class (oracle/make_golden.py)."""
    from emote_hack_b200.audio import SpeedEncoder
    gold = torch.load(GOLD / "speed_encoder.pt")
    enc = SpeedEncoder(9, 64)
    enc.load_state_dict(gold["state_dict"], strict=True)
    assert enc.bucket_centers == gold["centers"] and enc.bucket_radii == gold["radii"]
    out = enc.cuda()(gold["speeds"])
    assert out.shape == (9, 64) and rel_l2(out, gold["out"]) < 1e-5
    with pytest.raises(AssertionError):
        SpeedEncoder(10, 64)                                # the reference's own constructor check fails the same way
    with pytest.raises(AssertionError):
        enc(torch.zeros(2, 2))
