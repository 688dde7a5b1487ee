"""Importable alias of ``wav2vec-heart-sounds_b200/`` (a hyphen cannot appear in a Python module name).

``import wav2vec_heart_sounds_b200`` executes the real package's ``__init__`` with ``__path__`` pointed at the
hyphenated directory, so ``wav2vec_heart_sounds_b200.torchproc`` etc. resolve to the files there and the built
``libmpcg_b200.so`` stays in-tree beside them.
"""
import pathlib as _pathlib

_real = _pathlib.Path(__file__).resolve().parent.parent / "wav2vec-heart-sounds_b200"
__path__ = [str(_real)]
__file__ = str(_real / "__init__.py")
exec(compile((_real / "__init__.py").read_text(), __file__, "exec"))
