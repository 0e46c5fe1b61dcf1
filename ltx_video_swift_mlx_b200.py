"""Import shim: the package directory is named `ltx-video-swift-mlx_b200/` (not a valid Python identifier), so this
module gives it the importable name `ltx_video_swift_mlx_b200` by pointing `__path__` at that directory."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "ltx-video-swift-mlx_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
