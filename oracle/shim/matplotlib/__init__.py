cm = ticker = None
