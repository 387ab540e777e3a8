"""loss_functions.image_mse (reference loss_functions.py:1-5), kept as host code for the unfused path."""


def image_mse(mask, model_output, gt):
    if mask is None:
        return {'img_loss': ((model_output['model_out'] - gt['img']) ** 2).mean()}
    return {'img_loss': (mask * (model_output['model_out'] - gt['img']) ** 2).mean()}
