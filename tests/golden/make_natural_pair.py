#!/usr/bin/env python
"""Natural-image fixture for the parity tests: the reference's own sample pair (PytorchWCT/content/in4.jpg, style/in3.jpg,
both 512x512), stored as the original JPEG bytes so the repo carries 245 KB instead of decoded pixels.
Run in the build container (reads /root/reference); the GPU box only reads the committed npz."""
import os
import numpy as np
REF = "/root/reference/PytorchWCT"
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "natural_pair.npz")
c = np.frombuffer(open(os.path.join(REF, "content", "in4.jpg"), "rb").read(), dtype=np.uint8)
s = np.frombuffer(open(os.path.join(REF, "style", "in3.jpg"), "rb").read(), dtype=np.uint8)
np.savez(out, content_jpg=c, style_jpg=s, source=np.array(["PytorchWCT/content/in4.jpg", "PytorchWCT/style/in3.jpg"]))
print("wrote", out, c.size + s.size, "bytes")
