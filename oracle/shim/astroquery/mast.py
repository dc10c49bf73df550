Catalogs = Tesscut = None
