StatefulBrowser = None
