def __getattr__(name):
    raise AttributeError("matplotlib shim: pyplot." + name + " is not available in this image")
