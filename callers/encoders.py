"""Random-init stand-ins for the open_clip towers the reference trains from scratch (sparsify_clip.py:685-703:
`open_clip.create_model_and_transforms(config["model"], pretrained=None)`; open_clip is not installed here).

ViT-B/32 geometry for the image side (224^2 images, 32^2 patches -> 49 + 1 tokens, width 768, 12 layers, 12 heads)
and the CLIP text transformer (vocab 49408, 77 tokens, width 512, 12 layers, 8 heads, causal mask, features taken at
the highest token id = end of text), both projected to a 512-d joint space.  Attention is
F.scaled_dot_product_attention, the rest cuBLAS: the encoders are the load the loss is measured next to, not the
subject of this repo."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class Block(nn.Module):
    def __init__(self, width, heads, mlp_ratio=4):
        super().__init__()
        self.heads = heads
        self.ln_1 = nn.LayerNorm(width)
        self.qkv = nn.Linear(width, 3 * width)
        self.out = nn.Linear(width, width)
        self.ln_2 = nn.LayerNorm(width)
        self.fc = nn.Linear(width, mlp_ratio * width)
        self.proj = nn.Linear(mlp_ratio * width, width)

    def forward(self, x, causal=False):
        B, L, W = x.shape
        q, k, v = self.qkv(self.ln_1(x)).view(B, L, 3, self.heads, W // self.heads).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(q, k, v, is_causal=causal)
        x = x + self.out(a.transpose(1, 2).reshape(B, L, W))
        return x + self.proj(F.gelu(self.fc(self.ln_2(x))))


class VisionTower(nn.Module):
    def __init__(self, image_size=224, patch=32, width=768, layers=12, heads=12, out_dim=512):
        super().__init__()
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch, stride=patch, bias=False)
        n_tok = (image_size // patch) ** 2 + 1
        s = width ** -0.5
        self.class_embedding = nn.Parameter(s * torch.randn(width))
        self.positional_embedding = nn.Parameter(s * torch.randn(n_tok, width))
        self.ln_pre = nn.LayerNorm(width)
        self.blocks = nn.ModuleList([Block(width, heads) for _ in range(layers)])
        self.ln_post = nn.LayerNorm(width)
        self.proj = nn.Parameter(s * torch.randn(width, out_dim))

    def forward(self, images):
        x = self.conv1(images).flatten(2).transpose(1, 2)                      # [B, 49, W]
        cls = self.class_embedding.to(x.dtype).expand(x.shape[0], 1, -1)
        x = torch.cat([cls, x], dim=1) + self.positional_embedding.to(x.dtype)
        x = self.ln_pre(x)
        for b in self.blocks:
            x = b(x)
        return self.ln_post(x[:, 0]) @ self.proj


class TextTower(nn.Module):
    def __init__(self, vocab=49408, context=77, width=512, layers=12, heads=8, out_dim=512):
        super().__init__()
        self.token_embedding = nn.Embedding(vocab, width)
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        self.positional_embedding = nn.Parameter(0.01 * torch.randn(context, width))
        self.blocks = nn.ModuleList([Block(width, heads) for _ in range(layers)])
        self.ln_final = nn.LayerNorm(width)
        self.text_projection = nn.Parameter(width ** -0.5 * torch.randn(width, out_dim))

    def forward(self, tokens):
        x = self.token_embedding(tokens) + self.positional_embedding[:tokens.shape[1]]
        for b in self.blocks:
            x = b(x, causal=True)
        x = self.ln_final(x)
        return x[torch.arange(x.shape[0], device=x.device), tokens.argmax(dim=-1)] @ self.text_projection


class MiniCLIP(nn.Module):
    """encode_image / encode_text, the two methods the training loop calls (sparsify_clip.py:768-769)."""

    def __init__(self, vision=None, text=None):
        super().__init__()
        self.visual = VisionTower(**(vision or {}))
        self.text = TextTower(**(text or {}))

    def encode_image(self, images):
        return self.visual(images)

    def encode_text(self, tokens):
        return self.text(tokens)

    def forward(self, images, tokens):
        return self.encode_image(images), self.encode_text(tokens)
