"""Joint training of the demapper and the weighted decoder - drop-in for the reference's train_joint
(ofdm/ofdm_nn.py:257-396) on the native sparse backward (ldpc_bp_train_backward).

Same positional arguments, optimiser (SGD, demapper at 5x the learning rate, ofdm_nn.py:287-290), loss (BCE on
P(bit=1), gradient accumulated over the minibatches of a batch, ofdm_nn.py:327-343), per-epoch test print and
checkpoint dictionary (ofdm_nn.py:385-392).  Keyword-only extras: ``model_dir`` (the reference hard-codes
'outputs/model/'), ``minibatch_size`` (512 there), ``verbose``, ``return_model``.
The demapper-only trainers of that file (train_nn, train_nn_tanh) are outside the hot path.
"""
import collections
import datetime
import os

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim

from nn.joint import Joint

__all__ = ["train_joint"]


def train_joint(input_samples, output_samples, test_input, test_output, H, bp_iterations, clamp_value, data_timestamp, snrdb,
                learning_rate, qbits, clipdb, ofdm_size, num_epochs, batch_size, load_model=None, *, model_dir="outputs/model",
                minibatch_size=2 ** 9, verbose=True, return_model=False):
    snr = np.power(10, snrdb / 10)
    num_samples = input_samples.shape[0]
    num_batches = num_samples // batch_size
    minibatch_size = min(int(minibatch_size), int(batch_size))
    num_minibatches = batch_size // minibatch_size
    device = torch.device("cuda")                   # no CPU path
    model = nn.DataParallel(Joint(ofdm_size, snr, H, bp_iterations)) if torch.cuda.device_count() > 1 else _Wrapped(Joint(ofdm_size, snr, H, bp_iterations))
    model.to(device)
    criterion = nn.BCELoss()
    optimizer = optim.SGD([{"params": model.module.LLRest.parameters(), "lr": 5 * learning_rate},
                           {"params": model.module.BP.parameters()}], lr=learning_rate)
    if load_model:                                  # a demapper checkpoint: module.<layer> -> module.LLRest.<layer> (ofdm_nn.py:294-311)
        checkpoint = torch.load(os.path.join(model_dir, load_model), map_location=device, weights_only=False)
        d = collections.OrderedDict()
        for old_key, value in checkpoint["model_state_dict"].items():
            parts = old_key.split(".")
            parts.insert(1, "LLRest")
            d[".".join(parts)] = value
        model.load_state_dict(d, strict=False)

    E = model.module.layer_size()
    train_loss = np.zeros(num_epochs)
    epoch = -1
    for epoch in range(num_epochs):
        model.train()
        p = np.random.permutation(num_samples)      # shuffle every epoch
        input_samples, output_samples = input_samples[p], output_samples[p]
        for batch in range(num_batches):
            for mb in range(num_minibatches):
                a = batch * batch_size + mb * minibatch_size
                x_batch = torch.tensor(input_samples[a:a + minibatch_size], dtype=torch.float, requires_grad=True, device=device)
                y_batch = torch.tensor(output_samples[a:a + minibatch_size], dtype=torch.float, device=device)
                x_temp = torch.zeros(x_batch.shape[0], E, dtype=torch.float, device=device)
                loss = criterion(model(x_batch, x_temp, clamp_value), y_batch) / num_minibatches
                loss.backward()
                train_loss[epoch] += loss.item()
            optimizer.step()
            optimizer.zero_grad()
        with torch.no_grad():
            x_test = torch.tensor(test_input, dtype=torch.float, device=device)
            y_test = torch.tensor(test_output, dtype=torch.float, device=device)
            y_est = model(x_test, torch.zeros(x_test.shape[0], E, dtype=torch.float, device=device), clamp_value)
            test_loss = float(criterion(y_est, y_test))
        ber = float(np.mean(np.abs(np.round(y_est.cpu().numpy()) - np.round(test_output))))
        if verbose:
            print("[epoch %d] train_loss: %.3f, test_loss: %.3f, test_ber: %.3f" % (epoch + 1, train_loss[epoch] / max(num_batches, 1), test_loss, ber))

    ts = datetime.datetime.now()
    filename = ts.strftime("%Y%m%d-%H%M%S") + "_qbits={}_clipdb={}_snr={}_lr={}_joint.pth".format(qbits, clipdb, snrdb, learning_rate)
    os.makedirs(model_dir, exist_ok=True)
    torch.save({"epoch": epoch, "data_timestamp": data_timestamp, "batch_size": batch_size, "model_state_dict": model.state_dict(),
                "optimizer_state_dict": optimizer.state_dict(), "loss": train_loss}, os.path.join(model_dir, filename))
    return (filename, model) if return_model else filename


class _Wrapped(nn.Module):
    """Single-GPU stand-in for nn.DataParallel: same 'module.' prefix in the state_dict (the reference always wraps,
    ofdm_nn.py:280), no scatter/gather."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *args):
        return self.module(*args)
