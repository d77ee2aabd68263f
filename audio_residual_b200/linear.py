"""Mirror of the reference's src/linear.py: frozen CLAP encoder + trainable nn.Linear(512, n_classes) probe."""
import torch
import torch.nn.functional as F
from torch import nn


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
        return self.classifier(audio_embeds)


def train_linear_head_one_epoch(model, dataloader, optimizer, criterion, device):
    """src/linear.py:35-53"""
    model.train()
    total_loss, correct, total = 0.0, 0, 0
    for x, true_labels in dataloader:
        optimizer.zero_grad()
        logits = model(x, device)
        loss = criterion(logits, true_labels.to(device))
        loss.backward()
        optimizer.step()
        preds = logits.argmax(dim=-1).cpu()
        correct += (preds == true_labels).sum().item()
        total += x.size(0)
        total_loss += loss.item() * x.size(0)
    return total_loss / total, correct / total


def eval_linear_head(model, dataloader, device):
    """src/linear.py:97-125: (predictions, targets, softmax similarities)."""
    model.eval()
    all_preds, all_targets, all_similarities = [], [], []
    with torch.no_grad():
        for x, true_labels in dataloader:
            logits = model(x, device)
            all_preds.extend(logits.argmax(dim=-1).cpu().tolist())
            all_targets.extend(true_labels.tolist())
            all_similarities.append(F.softmax(logits, dim=-1))
    return all_preds, all_targets, torch.cat(all_similarities, dim=0).cpu().numpy()
