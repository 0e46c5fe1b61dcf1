"""Host-side mirror of the reference interface (no GPU): scheduler, shapes, validation, flag marshalling."""
import numpy as np
import pytest

from helpers import O, product

product()
from ltx_video_swift_mlx_b200 import latent_utils, scheduler  # noqa: E402
from ltx_video_swift_mlx_b200._lib import LtxError  # noqa: E402
from ltx_video_swift_mlx_b200.context import make_flags  # noqa: E402


@pytest.mark.parametrize("tokens", [None, 512, 1536, 4096, 6144, 50688])
def test_scheduler_mirror_matches_oracle_distilled(tokens):
    s = scheduler.LTXScheduler()
    got = s.set_timesteps(8, distilled=True, latent_token_count=tokens)
    assert np.allclose(got, O.set_timesteps(8, True, tokens), rtol=0, atol=1e-6)   # float32 evaluation order / expf differ by a few ulp between libms
    assert s.total_steps == 8 and s.initial_sigma == 1.0


@pytest.mark.parametrize("steps,tokens", [(40, 1536), (20, 512), (8, None), (2, 1536)])
def test_scheduler_mirror_matches_oracle_dev(steps, tokens):
    got = scheduler.LTXScheduler().set_timesteps(steps, False, tokens)
    assert np.allclose(got, O.set_timesteps(steps, False, tokens), rtol=0, atol=1e-6)
    assert got[0] == 1.0 and got[-1] == 0.0 and all(a > b for a, b in zip(got, got[1:]))


def test_custom_sigmas_and_constants():
    s = scheduler.LTXScheduler()
    s.set_custom_sigmas(scheduler.STAGE_2_DISTILLED_SIGMA_VALUES[:-1])
    assert s.sigmas == scheduler.STAGE_2_DISTILLED_SIGMA_VALUES
    assert scheduler.get_sigma_schedule(8, True) == scheduler.DISTILLED_SIGMA_VALUES


def test_shapes_and_validation():
    sh = latent_utils.VideoLatentShape.from_pixel_dimensions(1, 128, 25, 512, 768)
    assert sh.fhw == (4, 16, 24) and sh.token_count == 1536
    x = np.random.RandomState(0).randn(1, 128, 2, 3, 4).astype(np.float32)
    sh2 = latent_utils.VideoLatentShape(1, 128, 2, 3, 4)
    assert np.array_equal(latent_utils.unpatchify(latent_utils.patchify(x), sh2), x)
    latent_utils.LTXVideoGenerationConfig().validate()
    for bad in (dict(width=770), dict(num_frames=24), dict(num_frames=265), dict(num_steps=0), dict(cfg_scale=0.5)):
        with pytest.raises(LtxError) as e:
            latent_utils.LTXVideoGenerationConfig(**bad).validate()
        assert e.value.code == 1


def test_flag_marshalling():
    f = make_flags(stg_blocks=[29], skip_self_attn=True, cas_blocks=[1, 2], cross_attn_scale=0.5, context_key=9)
    assert f.n_stg_blocks == 1 and f.stg_blocks[0] == 29 and f.skip_self_attn == 1 and f.skip_ff == 0
    assert f.n_cas_blocks == 2 and f.cas_blocks[1] == 2 and abs(f.cross_attn_scale - 0.5) < 1e-7 and f.context_key == 9
