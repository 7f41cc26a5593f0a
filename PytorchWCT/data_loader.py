"""Content x style pairing and image decode (host I/O around the hot path; reference PytorchWCT/data_loader.py:21-59).
Same constructor and item contract: (content [3,H,W] in [0,1], style [3,H,W], "<content>+<style>.jpg")."""
import os

import torch
import torch.utils.data as data
import torchvision.transforms as transforms
from PIL import Image

IMG_EXT = (".png", ".jpg", ".jpeg")


def is_image_file(name):
    return name.endswith(IMG_EXT)


def default_loader(path):
    return Image.open(path).convert("RGB")


class Dataset(data.Dataset):
    def __init__(self, contentPath, stylePath, texturePath, c_size=0, s_size=0, picked_content_mark=".",
                 picked_style_mark=".", synthesis=False):
        super().__init__()
        self.content_size, self.style_size, self.synthesis = c_size, s_size, synthesis
        if synthesis:
            self.texturePath = texturePath
            self.items = [(None, t) for t in sorted(os.listdir(texturePath)) if is_image_file(t)]
        else:
            self.contentPath, self.stylePath = contentPath, stylePath
            cs = [x for x in os.listdir(contentPath) if is_image_file(x) and picked_content_mark in x]
            ss = [x for x in os.listdir(stylePath) if is_image_file(x) and picked_style_mark in x]
            self.items = [(c, s) for c in cs for s in ss]          # Cartesian pairing, content-major (data_loader.py:32-36)
        self.to_tensor = transforms.ToTensor()

    def __len__(self):
        return len(self.items)

    def _load(self, path, size):
        img = default_loader(path)
        if size:
            img = transforms.Resize(size)(img)                     # shorter side -> size (data_loader.py:52-55)
        return self.to_tensor(img)

    def paths(self, index):
        """(content path or None, style / texture path, output name) of pair `index` -- for the device-side decode
        (`--gpu_io`, collaborative_distillation_b200.image_io.load_image) that replaces __getitem__'s PIL pipeline."""
        c, s = self.items[index]
        if self.synthesis:
            return None, os.path.join(self.texturePath, s), s.split(".")[0] + ".jpg"
        return os.path.join(self.contentPath, c), os.path.join(self.stylePath, s), c.split(".")[0] + "+" + s.split(".")[0] + ".jpg"

    def __getitem__(self, index):
        c, s = self.items[index]
        if not self.synthesis:
            content = self._load(os.path.join(self.contentPath, c), self.content_size)
            style = self._load(os.path.join(self.stylePath, s), self.style_size)
            return content, style, c.split(".")[0] + "+" + s.split(".")[0] + ".jpg"
        # texture synthesis: noise content of the texture's size (the reference's rand_like(PIL) at :74 cannot run)
        tex = default_loader(os.path.join(self.texturePath, s))
        if self.style_size:
            w, h = tex.size
            tex = tex.resize((self.style_size, int(h * self.style_size / w)) if w > h else (int(w * self.style_size / h), self.style_size))
        tex = self.to_tensor(tex)
        return torch.rand_like(tex), tex, s.split(".")[0] + ".jpg"
