"""Import helper for packages the reference lists but this image lacks (imageio)."""
import os
import sys


def imageio_v2():
    try:
        import imageio.v2 as iio
    except ImportError:
        shim = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_shims")
        if shim not in sys.path:
            sys.path.append(shim)
        import imageio.v2 as iio
    return iio
