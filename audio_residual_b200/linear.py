"""Mirror of the reference's src/linear.py: frozen CLAP encoder + trainable nn.Linear(512, n_classes) probe.
HTSATLinearClassifier :9-32, train_linear_head_one_epoch :35-53, train_and_eval_linear_head :56-95, eval_linear_head :97-125.
The probe's matmul, the cross entropy and their backward run in libard_b200.so (`ard_head_forward/backward`, `ard_ce_forward`)."""
import gc
import os

import numpy as np
import torch
from torch import nn

from .head import apply_criterion, head_logits


class HTSATLinearClassifier(nn.Module):
    """src/linear.py:9-32"""

    def __init__(self, clap, n_classes, feat_dim=512):
        super().__init__()
        self.clap = clap
        self.feat_dim = feat_dim
        self.n_classes = n_classes
        for p in self.clap.parameters():
            p.requires_grad = False
        self.classifier = nn.Linear(self.feat_dim, self.n_classes)
        nn.init.kaiming_normal_(self.classifier.weight)
        nn.init.zeros_(self.classifier.bias)

    def forward(self, x, device):
        audio_data = x.squeeze(1).to(device)
        audio_embeds = self.clap.get_audio_embedding_from_data(x=audio_data, use_tensor=True)
        audio_embeds = audio_embeds.to(device).float()
        return head_logits(audio_embeds, self.classifier.weight, self.classifier.bias)


def train_linear_head_one_epoch(model, dataloader, optimizer, criterion, device):
    """src/linear.py:35-53"""
    model.train()
    total_loss, correct, total = 0.0, 0, 0
    for x, true_labels in dataloader:
        optimizer.zero_grad()
        logits = model(x, device)
        loss = apply_criterion(criterion, logits, true_labels.to(device))
        loss.backward()
        optimizer.step()
        preds = logits.argmax(dim=-1).cpu()
        correct += (preds == true_labels).sum().item()
        total += x.size(0)
        total_loss += loss.item() * x.size(0)
    return total_loss / total, correct / total


def train_and_eval_linear_head(clap, dataset_name, folds, n_classes, save_dir, lr=0.01, epochs=10):
    """src/linear.py:56-95: K-fold linear-probe training (AdamW) and per-fold .npz of predictions / targets / softmax scores."""
    save_dir = os.path.join(save_dir, dataset_name, "Linear")
    os.makedirs(save_dir, exist_ok=True)
    device = clap.device
    for i, (train_load, val_load) in enumerate(folds):
        print(f"===== Eval fold {i} =====")
        save_file = os.path.join(save_dir, f"evalfold_{i}.npz")
        model = HTSATLinearClassifier(clap=clap, n_classes=n_classes).to(device)
        optimizer = torch.optim.AdamW(filter(lambda p: p.requires_grad, model.parameters()), lr=lr)
        criterion = nn.CrossEntropyLoss()
        for ep in range(epochs):
            print(f"=== Epoch {ep} ===")
            loss, acc = train_linear_head_one_epoch(model, train_load, optimizer, criterion, device)
            print(f"Train loss: {loss}, Train accuracy: {acc}")
        preds, targs, similarities = eval_linear_head(model, val_load, device)
        np.savez_compressed(save_file, similarities=similarities, predictions=np.array(preds), targets=np.array(targs))
        torch.cuda.empty_cache()
        gc.collect()


def eval_linear_head(model, dataloader, device):
    """src/linear.py:97-125: (predictions, targets, softmax scores)."""
    model.eval()
    all_preds, all_targets, all_similarities = [], [], []
    with torch.no_grad():
        for x, true_labels in dataloader:
            logits = model(x, device)
            all_preds.extend(logits.argmax(dim=-1).cpu().tolist())
            all_targets.extend(true_labels.tolist())
            all_similarities.append(torch.softmax(logits, dim=-1))
    return all_preds, all_targets, torch.cat(all_similarities, dim=0).cpu().numpy()
