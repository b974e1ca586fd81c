import numpy as np
from PIL import Image


def imwrite(path, array, **_):
    Image.fromarray(np.asarray(array)).save(path)


def imread(path, **_):
    return np.asarray(Image.open(path))


def mimsave(path, frames, fps=10, loop=0, duration=None, **_):
    ims = [Image.fromarray(np.asarray(f)) for f in frames]
    if not ims:
        raise ValueError("mimsave: no frames")
    ms = int(round(duration * 1000)) if duration else int(round(1000.0 / fps))
    ims[0].save(path, save_all=True, append_images=ims[1:], duration=ms, loop=loop)
