"""Multi-head self-attention layer of the Transformer variants
(/root/reference/pytorch/models.py:587-665).  NOT IMPLEMENTED YET in this round: the entry points
fail loudly rather than falling back to PyTorch."""


def _todo():
    raise NotImplementedError(
        'Cnn_9layers_Transformer_* / MultiHead: the attention kernels are not built yet '
        '(DESIGN.md section 7, next rows); there is deliberately no PyTorch fallback')


def multihead_forward(mh, feat, training, keep):
    _todo()


def multihead_backward(mh, ctx, dfeat, grad_of):
    _todo()


def multihead_module_forward(mh, q, k, v, mask=None):
    _todo()
