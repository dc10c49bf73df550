BeautifulSoup = None
