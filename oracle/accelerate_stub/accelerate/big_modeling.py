def dispatch_model(model, *a, **k):
    return model


def infer_auto_device_map(*a, **k):
    return {}
