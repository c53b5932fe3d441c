"""
fib_tf_b200.screen -- optional, headless stand-in for the reference's SDL2 window (screen.py).

The reference's Screen opens an SDL2 window through ctypes and needs libSDL2, libSDL2_ttf and
matplotlib at import time (screen.py:10-22).  Visualisation is out of scope for the hot path;
what run(im) needs from `im` is only imshow(image) and wait() (ionic.py:206-245).  This class keeps
that protocol (plus peek/plot/draw_text/save as used by the reference drivers) and records frames
instead of painting them, so drivers written for a Screen run unchanged on a GPU box.
"""
import numpy as np


class Screen:
    def __init__(self, height, width, caption='', keep_every=0):
        self.height, self.width, self.caption = height, width, caption
        self.frames_shown = 0
        self.last = None
        self.keep_every = keep_every      # > 0: keep every n-th frame in self.frames
        self.frames = []
        self.texts = []

    def imshow(self, image):
        """Receives a [height, width] float frame in 0..1 (run() passes image()*phase)."""
        a = np.asarray(image, dtype=np.float32)
        if a.shape != (self.height, self.width):
            raise ValueError('frame shape %r != screen %r' % (a.shape, (self.height, self.width)))
        self.last = a
        if self.keep_every and self.frames_shown % self.keep_every == 0:
            self.frames.append(a.copy())
        self.frames_shown += 1
        return True

    def peek(self):
        """The reference polls SDL events here and returns False on quit; headless: always True."""
        return True

    def wait(self):
        """The reference blocks until the window is closed; headless: returns at once."""
        return None

    def plot(self, *args, **kw):
        return None

    def draw_text(self, text, *args, **kw):
        self.texts.append(str(text))

    def save(self, name):
        """Saves the last frame: PNG through Pillow when it is installed and the name asks for it,
        otherwise .npy."""
        if self.last is None:
            raise RuntimeError('no frame has been shown yet')
        if str(name).lower().endswith('.png'):
            try:
                from PIL import Image
                Image.fromarray((np.clip(self.last, 0, 1) * 255).astype(np.uint8)).save(name)
                return name
            except ImportError:
                name = str(name)[:-4] + '.npy'
        np.save(name, self.last)
        return name
