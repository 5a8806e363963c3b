"""pytorch_utils drop-in (/root/reference/pytorch/pytorch_utils.py): move_data_to_device,
append_to_dict, forward (inference loop), do_mixup."""
import numpy as np
import torch


def move_data_to_device(x, device):
    if 'float' in str(x.dtype):
        x = torch.Tensor(x)
    elif 'int' in str(x.dtype):
        x = torch.LongTensor(x)
    else:
        return x
    return x.to(device)


def append_to_dict(dict, key, value):
    dict.setdefault(key, []).append(value)


def do_mixup(x, mixup_lambda):
    """out[i] = x[2i] * lam[2i] + x[2i+1] * lam[2i+1] (pytorch_utils.py:80-93).  Inside the models the
    feature mixup is fused into the bn0 kernel; this host-visible version serves the *targets*
    (main.py:246): a (2B, 17) tensor, i.e. plumbing-sized."""
    lam = mixup_lambda.to(x.dtype)
    shape = (-1,) + (1,) * (x.dim() - 1)
    return x[0::2] * lam[0::2].reshape(shape) + x[1::2] * lam[1::2].reshape(shape)


def forward(model, data_loader, return_input=False, return_target=False):
    """Inference loop of pytorch_utils.py:25-77: eval mode, no grad, per-batch outputs gathered on
    the host as numpy arrays."""
    device = next(model.parameters()).device
    output_dict = {}
    for batch_data_dict in data_loader:
        batch_waveform = move_data_to_device(batch_data_dict['waveform'], device)
        with torch.no_grad():
            model.eval()
            batch_output = model(batch_waveform)
        append_to_dict(output_dict, 'audio_name', batch_data_dict['audio_name'])
        append_to_dict(output_dict, 'clipwise_output', batch_output['clipwise_output'].data.cpu().numpy())
        if 'framewise_output' in batch_output.keys():
            append_to_dict(output_dict, 'framewise_output', batch_output['framewise_output'].data.cpu().numpy())
        if return_input:
            append_to_dict(output_dict, 'waveform', batch_data_dict['waveform'])
        if return_target:
            for key in ('target', 'strong_target'):
                if key in batch_data_dict.keys():
                    append_to_dict(output_dict, key, batch_data_dict[key])
    for key in output_dict.keys():
        output_dict[key] = np.concatenate(output_dict[key], axis=0)
    return output_dict
