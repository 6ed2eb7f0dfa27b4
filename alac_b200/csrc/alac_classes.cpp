// placeholder: ALACEncoder / ALACDecoder class layer (filled in next)
