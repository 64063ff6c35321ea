time_dilation = early_done = monitor = agents = augmentation = None
