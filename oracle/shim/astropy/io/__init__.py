fits = None
