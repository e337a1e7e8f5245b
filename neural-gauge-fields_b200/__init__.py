"""ngf_b200 — B200 (sm_100a) implementation of the volumetric-rendering hot path of fnzhan/Neural-Gauge-Fields.

Layout (only what the path needs):
  csrc/         hand-written CUDA kernels + the C ABI (``include/ngf_b200.h``) -> ``libngf_b200.so``
  _lib.py       ctypes binding of the C ABI
  field_base.py / triplane.py / infoinv.py / networks.py
                host-side mirrors of the reference's model classes (same names, ctor, forward signature)
  neutex.py     host-side mirror of the UV-Mapping ``NeuTex`` module (render path)
  render.py     ``renderer`` (reference: TriPlane/main.py:60-71) + multi-GPU ray sharding
  synth.py      seeded synthetic cameras / fields used by tests and bench (no dataset exists offline)

The package directory is called ``neural-gauge-fields_b200``; ``ngf_b200.py`` at the repo root makes it importable.
"""
from . import _lib  # noqa: F401
from .field_base import AlphaGridMask, Base  # noqa: F401
from .triplane import TriPlane  # noqa: F401
from .infoinv import TriPlane as InfoInvTriPlane  # noqa: F401
from .neutex import NeuTex  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .render import FrameComm, ShardedFrameRenderer, frame_post, renderer, render_rays, render_frames, render_frame_sharded, shard_rays, unshard_frame, visualize_depth  # noqa: F401

__version__ = "0.1.0"
