"""Minimal stand-in for the `imageio` package (absent from this image; the reference's scripts do
`import imageio.v2 as imageio` -- src/train.py:5, src/main.py:3, src/make_gif.py:1).  Only the two
calls those scripts make are provided, on top of Pillow.  It is put on sys.path only when the real
package cannot be imported."""
from . import v2  # noqa: F401
from .v2 import imread, imwrite, mimsave  # noqa: F401
